"""Multi-GPU parity check, launched as one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Every rank runs the partitioned static step (NCCL halo sums + all-reduced dot products); rank 0 also solves the whole
mesh on its own GPU and compares displacements / reaction forces (<= 1e-10 relative) and CG iteration counts."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from en234_fea_b200 import GpuSystem, Model, Partition, synthetic_block_input  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    only = None                                   # --dynamics-only / --project-only: one section (GPU-minute budget)
    for flag in ("--dynamics-only", "--project-only", "--c3d8-only", "--static-only", "--print-only", "--latency-only"):
        if flag in sys.argv:
            sys.argv.remove(flag)
            only = flag[2:-5]
    dims = tuple(int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (12, 10, 8 * world)))
    jitter = float(sys.argv[4]) if len(sys.argv) >= 5 else 0.2      # 0 with power-of-two dims: congruent-mesh operator
    ok = True
    for method in ((0, 1) if only in (None, "static") else ()):
        for load in ("traction", "stretch"):
            m = Model.from_text(synthetic_block_input(*dims, jitter=jitter, cg_tolerance=1e-26, cg_maxit=20000, load=load,
                                                      steps=2 if load == "stretch" else 1))
            part = Partition(m, world, rank)
            g = GpuSystem(m, device=local, part=part)
            comm_kind = g.setup_distributed()
            res = g.run_static(method=method, check_linear_cg=True)
            l2g = part.ints("local_to_global_node").astype(np.int64)
            ldof = (3 * l2g[:, None] + np.arange(3)).ravel()
            # gather the distributed fields on rank 0 (owned entries win; shared copies must agree)
            ug = torch.zeros(3 * m.n_nodes, dtype=torch.float64, device="cuda")
            rg = torch.zeros_like(ug)
            cnt = torch.zeros_like(ug)
            idx = torch.from_numpy(ldof).cuda()
            ug[idx] = torch.from_numpy(res["dof_total"]).cuda()
            rg[idx] = torch.from_numpy(res["rforce"]).cuda()
            cnt[idx] = 1.0
            dist.all_reduce(ug); dist.all_reduce(rg); dist.all_reduce(cnt)
            ug, rg = (ug / cnt).cpu().numpy(), (rg / cnt).cpu().numpy()
            # shared copies agree bit for bit => dividing the sum by the count reproduces each copy
            agree = np.array_equal(ug[ldof], res["dof_total"]) or np.abs(ug[ldof] - res["dof_total"]).max() <= 1e-15 * np.abs(ug).max()
            if rank == 0:
                g1 = GpuSystem(m, device=local)
                ref = g1.run_static(method=0, check_linear_cg=True)
                eu = np.abs(ug - ref["dof_total"]).max() / np.abs(ref["dof_total"]).max()
                er = np.abs(rg - ref["rforce"]).max() / np.abs(ref["rforce"]).max()
                its, its1 = res["cg_iterations"].tolist(), ref["cg_iterations"].tolist()
                good = eu < 1e-10 and er < 1e-10 and all(abs(a - b) <= max(2, b // 25) for a, b in zip(its, its1)) and \
                    res["newton"].shape == ref["newton"].shape and agree
                print("mgpu_check [%s] world=%d mesh=%s jitter=%g congruent=%d method=%d load=%s: u err %.2e, rforce err %.2e, CG its %s vs %s, copies agree %s -> %s"
                      % (comm_kind, world, dims, jitter, g.stats().mf_congruent, method, load, eu, er, its, its1, agree, "OK" if good else "FAIL"), flush=True)
                ok = ok and good
                g1.close()
            dist.barrier()
            g.close()
    # explicit dynamic step on the partitioned mesh: one halo sum of the element forces per step (and of the lumped mass once)
    if only in (None, "dynamics"):
        ok = dynamics_section(rank, world, local, dims, jitter) and ok
    if only in (None, "project"):
        ok = project_section(rank, world, local, dims, jitter) and ok
    if only in (None, "c3d8"):
        ok = c3d8_section(rank, world, local) and ok
    if only in (None, "print"):
        ok = print_section(rank, world, local, dims, jitter) and ok
    if only == "latency":
        latency_section(rank, world, local, dims, jitter)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


def dynamics_section(rank, world, local, dims, jitter):
    ok = True
    text = synthetic_block_input(*dims, jitter=jitter, element="U1", props=[100.0, 0.3, 2.5], load="stretch", steps=4, dt=0.5,
                                 dynamic=(1e-3 * 8.0 / max(dims), 70))
    text = text.replace(" END DEGREES OF FREEDOM", " END DEGREES OF FREEDOM\n FORCES\n  xmax, 3, HISTORY, ramp\n END FORCES")
    m = Model.from_text(text)
    part = Partition(m, world, rank)
    g = GpuSystem(m, device=local, part=part)
    comm_kind = g.setup_distributed()
    res = g.run_dynamic()
    l2g = part.ints("local_to_global_node").astype(np.int64)
    ldof = (3 * l2g[:, None] + np.arange(3)).ravel()
    idx = torch.from_numpy(ldof).cuda()
    fields = {}
    for k in ("dof_total", "velocity", "lumped_mass"):
        acc = torch.zeros(3 * m.n_nodes, dtype=torch.float64, device="cuda")
        cnt = torch.zeros_like(acc)
        acc[idx] = torch.from_numpy(res[k]).cuda()
        cnt[idx] = 1.0
        dist.all_reduce(acc); dist.all_reduce(cnt)
        fields[k] = (acc / cnt).cpu().numpy()
    agree = np.abs(fields["dof_total"][ldof] - res["dof_total"]).max() <= 1e-15 * np.abs(fields["dof_total"]).max()
    if rank == 0:
        g1 = GpuSystem(m, device=local)
        ref = g1.run_dynamic()
        errs = {k: np.abs(fields[k] - ref[k]).max() / np.abs(ref[k]).max() for k in fields}
        good = all(e < 1e-10 for e in errs.values()) and agree and res["steps"] == ref["steps"] == 70
        print("mgpu_check [%s] world=%d mesh=%s explicit dynamics 70 steps: u err %.2e, v err %.2e, mass err %.2e, copies agree %s, "
              "%.1f us/step on %d GPUs vs %.1f us/step on one -> %s"
              % (comm_kind, world, dims, errs["dof_total"], errs["velocity"], errs["lumped_mass"], agree,
                 1e3 * res["device_ms"] / 70, world, 1e3 * ref["device_ms"] / 70, "OK" if good else "FAIL"), flush=True)
        ok = ok and good
        g1.close()
    dist.barrier()
    g.close()
    return ok


def c3d8_section(rank, world, local):
    """The reference's golden C3D8 + UMAT model (unsymmetric tangent -> Jacobi-BiCGSTAB with halo sums and all-reduced dot
    products) on an element partition: same Newton history as one GPU, displacements within 1e-8 (the golden log's own
    records are reproduced to their printed digits by the one-GPU path, tests/test_gpu_parity.py)."""
    import gzip
    golden = os.path.join(ROOT, "tests", "golden", "holeplate_3d_umat.in.gz")
    m = Model.from_text(gzip.open(golden, "rb").read().decode())
    part = Partition(m, world, rank)
    g = GpuSystem(m, device=local, part=part)
    comm_kind = g.setup_distributed()
    import time
    t0 = time.perf_counter()
    res = g.run_static(method=0)
    t_part = time.perf_counter() - t0
    st = g.stats()
    last_part = 1e3 * st.solve_ms / max(1, st.last_cg_iterations)
    l2g = part.ints("local_to_global_node").astype(np.int64)
    ldof = (3 * l2g[:, None] + np.arange(3)).ravel()
    acc = torch.zeros(3 * m.n_nodes, dtype=torch.float64, device="cuda")
    cnt = torch.zeros_like(acc)
    idx = torch.from_numpy(ldof).cuda()
    acc[idx] = torch.from_numpy(res["dof_total"]).cuda()
    cnt[idx] = 1.0
    dist.all_reduce(acc); dist.all_reduce(cnt)
    ug = (acc / cnt).cpu().numpy()
    ok = True
    if rank == 0:
        g1 = GpuSystem(m, device=local)
        t0 = time.perf_counter()
        ref = g1.run_static(method=0)
        t_one = time.perf_counter() - t0
        st1 = g1.stats()
        print("   last solve: %.1f us per BiCGSTAB iteration on %d GPUs (device time), %.1f on one" %
              (last_part, world, 1e3 * st1.solve_ms / max(1, st1.last_cg_iterations)), flush=True)
        eu = np.abs(ug - ref["dof_total"]).max() / np.abs(ref["dof_total"]).max()
        same = res["newton"].shape == ref["newton"].shape
        rel = np.abs(res["newton"][:, 2:5] - ref["newton"][:, 2:5]).max() / np.abs(ref["newton"][:, 2:5]).max() if same else np.inf
        good = same and eu < 1e-8 and rel < 1e-5
        print("mgpu_check [%s] world=%d golden C3D8 + UMAT hole plate (BiCGSTAB on partitions): %d Newton records, norms agree to %.1e, "
              "u err %.2e, BiCGSTAB its %s vs %s, %.0f us per iteration on %d GPUs vs %.0f on one (whole step / iterations) -> %s"
              % (comm_kind, world, res["newton"].shape[0], rel, eu, res["cg_iterations"].tolist(), ref["cg_iterations"].tolist(),
                 1e6 * t_part / max(1, int(res["cg_iterations"].sum())), world, 1e6 * t_one / max(1, int(ref["cg_iterations"].sum())),
                 "OK" if good else "FAIL"), flush=True)
        ok = good
        g1.close()
    dist.barrier()
    g.close()
    return ok


def latency_section(rank, world, local, dims, jitter):
    """Collectives of the PCG loop in isolation (en234gpu_bench_collective): us per all-reduce kernel, per halo sum, per empty
    kernel, back to back inside a captured graph on every rank at once."""
    import ctypes as C
    from en234_fea_b200 import gpu_lib
    m = Model.from_text(synthetic_block_input(*dims, jitter=jitter))
    part = Partition(m, world, rank)
    g = GpuSystem(m, device=local, part=part)
    comm_kind = g.setup_distributed()
    out = []
    for which in (2, 0, 1):
        us = C.c_double()
        dist.barrier()
        if gpu_lib().en234gpu_bench_collective(g.h, C.c_int(which), C.c_int(20), C.byref(us)):
            raise RuntimeError(gpu_lib().en234gpu_last_error().decode())
        t = torch.tensor([us.value], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(float(t.item()))
    if rank == 0:
        print("mgpu_check [%s] world=%d mesh=%s collectives in isolation (max over ranks, 1280 back-to-back calls in graphs): empty kernel %.2f us, "
              "scalar all-reduce kernel %.2f us, halo sum of %d interface nodes (push + wait) %.2f us"
              % (comm_kind, world, dims, out[0], out[1], part.ints("shared_nodes").size, out[2]), flush=True)
    dist.barrier()
    g.close()


def print_section(rank, world, local, dims, jitter):
    """PRINT STATE on a partitioned run: every rank projects the fields on its part, rank 0 gathers them over the communicator
    and writes ONE Tecplot file; compared number by number with the file of the same analysis on one GPU."""
    import tempfile
    names = ["S11", "S22", "S33", "S12", "S13", "S23", "SMISES"]
    tmp = tempfile.gettempdir()
    paths = [os.path.join(tmp, "en234_mgpu_state_%d_%s.dat" % (world, k)) for k in ("part", "one")]
    m = Model.from_text(synthetic_block_input(*dims, jitter=jitter, cg_tolerance=1e-26, cg_maxit=20000, load="stretch", steps=2))
    m.set_state_print(paths[0], field_variables=names, print_dof=True, displaced_mesh=True, displacement_scale_factor=2.0)
    part = Partition(m, world, rank)
    g = GpuSystem(m, device=local, part=part)
    comm_kind = g.setup_distributed()
    res = g.run_static(method=0)
    ok = True
    if rank == 0:
        m.set_state_print(paths[1], field_variables=names, print_dof=True, displaced_mesh=True, displacement_scale_factor=2.0)
        g1 = GpuSystem(m, device=local)
        ref = g1.run_static(method=0)
        a, b = open(paths[0]).read().splitlines(), open(paths[1]).read().splitlines()
        same_shape = len(a) == len(b) and res["n_state_prints"] == ref["n_state_prints"] == 2
        worst, text_same = 0.0, True
        if same_shape:
            ncol = 3 + 6 + len(names)
            scale = None
            rows_a, rows_b = [], []
            for la, lb in zip(a, b):
                if len(la) == 19 * ncol and la[1:19].strip()[0] in "-0123456789":
                    rows_a.append([float(la[19 * k:19 * k + 19]) for k in range(ncol)])
                    rows_b.append([float(lb[19 * k:19 * k + 19]) for k in range(ncol)])
                else:
                    text_same = text_same and la == lb
            A, B = np.array(rows_a), np.array(rows_b)
            scale = np.abs(B).max(axis=0)
            worst = float((np.abs(A - B).max(axis=0) / np.where(scale > 0, scale, 1.0)).max())
            same_shape = same_shape and A.shape[0] == 2 * m.n_nodes
        good = same_shape and text_same and worst < 1e-9
        print("mgpu_check [%s] world=%d mesh=%s PRINT STATE on the partitioned run: %d zones, %d lines, headers / connectivity identical %s, "
              "worst column difference %.1e of the column maximum -> %s" % (comm_kind, world, dims, res["n_state_prints"], len(a), text_same, worst,
                                                                            "OK" if good else "FAIL"), flush=True)
        ok = good
        g1.close()
        for q in paths:
            if os.path.exists(q):
                os.remove(q)
    dist.barrier()
    g.close()
    return ok


def project_section(rank, world, local, dims, jitter):
    # state print projection on the partitioned mesh (rank-local numbering), consistent and lumped, against one GPU
    ok = True
    m = Model.from_text(synthetic_block_input(*dims, jitter=jitter))
    part = Partition(m, world, rank)
    g = GpuSystem(m, device=local, part=part)
    comm_kind = g.setup_distributed()
    l2g = part.ints("local_to_global_node").astype(np.int64)
    ldof = (3 * l2g[:, None] + np.arange(3)).ravel()
    ug = 0.01 * np.random.default_rng(7).uniform(-1, 1, 3 * m.n_nodes)
    codes = [0, 1, 2, 3, 4, 5, 6]
    proj = {}
    for lumped in (False, True):
        f = g.project_field_codes(codes, ug[ldof], lumped=lumped)
        acc = torch.zeros((m.n_nodes, len(codes)), dtype=torch.float64, device="cuda")
        cnt = torch.zeros(m.n_nodes, dtype=torch.float64, device="cuda")
        nidx = torch.from_numpy(l2g).cuda()
        acc[nidx] = torch.from_numpy(f).cuda()
        cnt[nidx] = 1.0
        dist.all_reduce(acc); dist.all_reduce(cnt)
        proj[lumped] = ((acc / cnt[:, None]).cpu().numpy(), np.abs((acc / cnt[:, None]).cpu().numpy()[l2g] - f).max())
    if rank == 0:
        g1 = GpuSystem(m, device=local)
        good = True
        msg = []
        for lumped in (False, True):
            ref = g1.project_field_codes(codes, ug, lumped=lumped)
            err = np.abs(proj[lumped][0] - ref).max() / np.abs(ref).max()
            good = good and err < 1e-10 and proj[lumped][1] <= 1e-12 * np.abs(ref).max()
            msg.append("%s err %.2e (copies differ by %.1e)" % ("lumped" if lumped else "consistent", err, proj[lumped][1]))
        print("mgpu_check [%s] world=%d mesh=%s state-print projection of 7 fields: %s -> %s"
              % (comm_kind, world, dims, ", ".join(msg), "OK" if good else "FAIL"), flush=True)
        ok = ok and good
        g1.close()
    dist.barrier()
    g.close()
    return ok


if __name__ == "__main__":
    main()
