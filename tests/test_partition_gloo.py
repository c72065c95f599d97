"""CPU test of the N > 1 host logic (world_size 2, gloo): the element partition produced by the product's host
library (en234host_partition), the halo-sum protocol (send partials to each neighbour in the shared-node order, add what
is received) and the ownership mask used for dot products.  The local operator on each rank is the ORACLE's assembled
matrix of the rank's sub-mesh (tests may use the oracle); the distributed result must equal the global oracle result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_binding as ob
from en234_fea_b200 import Model, Partition, synthetic_block_input


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dims, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = Model.from_text(synthetic_block_input(*dims, jitter=0.2))
        a = m.arrays()
        part = Partition(m, world, rank)
        l2g = part.ints("local_to_global_node").astype(np.int64)
        lel = part.ints("local_elements")
        owned = part.owned().astype(bool)
        # rank-local model for the oracle: local nodes / elements, no boundary conditions
        la = {"n_nodes": np.array([l2g.size], np.int32), "n_elements": np.array([lel.size], np.int32),
              "coords": a["coords"].reshape(-1, 3)[l2g].ravel(), "connectivity": part.ints("connectivity"),
              "element_flag": a["element_flag"][lel], "element_prop_index": a["element_prop_index"][lel],
              "element_n_props": a["element_n_props"][lel], "element_properties": a["element_properties"]}
        om = ob.OracleModel(la)
        s = ob.OracleSystem(om, 2)
        nl = 3 * l2g.size
        s.assemble(np.zeros(nl), np.zeros(nl))
        A_loc, _ = s.matrix()
        ldof = (3 * l2g[:, None] + np.arange(3)).ravel()
        rng = np.random.default_rng(42)
        xg = rng.standard_normal(3 * m.n_nodes)           # same on every rank
        y = A_loc @ xg[ldof]                              # partial sums on interface nodes
        # ---- halo sum exactly as the device path does it: pack per neighbour, exchange, add in neighbour order
        nbr, sp, sn = part.ints("neighbour_rank"), part.ints("shared_ptr"), part.ints("shared_nodes")
        send = [torch.from_numpy(y.reshape(-1, 3)[sn[sp[k]:sp[k + 1]]].copy()) for k in range(nbr.size)]
        recv = [torch.zeros_like(t) for t in send]
        reqs = []
        for k, r in enumerate(nbr.tolist()):
            reqs.append(dist.isend(send[k], dst=r))
            reqs.append(dist.irecv(recv[k], src=r))
        for r in reqs:
            r.wait()
        y = y.reshape(-1, 3)
        for k in range(nbr.size):
            y[sn[sp[k]:sp[k + 1]]] += recv[k].numpy()
        y = y.ravel()
        # ---- owned-only dot product, all-reduced
        dmask = np.repeat(owned, 3)
        dot = torch.tensor([float(xg[ldof][dmask] @ y[dmask])], dtype=torch.float64)
        dist.all_reduce(dot)
        # p.Ap from partial products over ALL local rows (no halo needed): must equal the owned-only value
        dot2 = torch.tensor([float(xg[ldof] @ (A_loc @ xg[ldof]))], dtype=torch.float64)
        dist.all_reduce(dot2)
        # ---- global reference on rank 0
        ok, msg = True, ""
        if rank == 0:
            og = ob.OracleModel({k: a[k] for k in ("n_nodes", "n_elements", "coords", "connectivity", "element_flag",
                                                   "element_prop_index", "element_n_props", "element_properties")})
            sg = ob.OracleSystem(og, 2)
            sg.assemble(np.zeros(3 * m.n_nodes), np.zeros(3 * m.n_nodes))
            A, _ = sg.matrix()
            yg = A @ xg
            e1 = np.abs(y - yg[ldof]).max() / np.abs(yg).max()
            e2 = abs(dot.item() - xg @ yg) / abs(xg @ yg)
            e3 = abs(dot2.item() - xg @ yg) / abs(xg @ yg)
            ok = e1 < 1e-13 and e2 < 1e-13 and e3 < 1e-13
            msg = "halo %.2e dot(owned) %.2e dot(partial) %.2e" % (e1, e2, e3)
        else:
            og = ob.OracleModel({k: a[k] for k in ("n_nodes", "n_elements", "coords", "connectivity", "element_flag",
                                                   "element_prop_index", "element_n_props", "element_properties")})
            sg = ob.OracleSystem(og, 2)
            sg.assemble(np.zeros(3 * m.n_nodes), np.zeros(3 * m.n_nodes))
            A, _ = sg.matrix()
            e1 = np.abs(y - (A @ xg)[ldof]).max() / np.abs(A @ xg).max()
            ok = e1 < 1e-13
            msg = "halo %.2e" % e1
        q.put((rank, ok, msg))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dims", [(4, 3, 6), (5, 5, 2)])
def test_two_rank_halo_sum_and_ownership(dims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dims, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, (rank, msg)
