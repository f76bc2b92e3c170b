"""N>1 logic on the CPU (gloo, world_size 2): instance sharding of the batched driver + the single gather of
step results.  The per-rank solver here is the oracle's C restatement (the CUDA engine needs a GPU); what is
under test is the host-side partition / schedule / gather plumbing that bench.py uses under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grid2op_b200.gridmodel import GridModel
from grid2op_b200.rollout import instance_schedule

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rows_to_inj(gm, rows):
    sl = gm.inj_slices()
    nl, ng = gm.n_load, gm.n_gen
    inj = np.tile(gm.default_inj(), (rows.shape[0], 1))
    inj[:, sl["load_p"]] = rows[:, :nl]
    inj[:, sl["load_q"]] = rows[:, nl:2 * nl]
    inj[:, sl["gen_p"]] = rows[:, 2 * nl:2 * nl + ng]
    inj[:, sl["gen_vm"]] = (rows[:, 2 * nl + ng:] / gm.prod_pu_to_kv[None, :]).astype(np.float32)
    return inj


def _solve(gm, chron, scen, t0, step):
    from oracle.c_oracle import COracle
    rows = chron[scen, (t0 + step) % chron.shape[1]]
    out, status, _, _ = COracle(gm, nthreads=1).run(np.tile(gm.default_topo(), (len(scen), 1)), _rows_to_inj(gm, rows))
    assert (status == 0).all()
    a_or = out[:, 3 * gm.n_line:4 * gm.n_line]
    return (a_or / gm.thermal_limit_a[None, :]).astype(np.float32)


def _worker(rank, world, port, per_rank, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    scen, t0 = instance_schedule(per_rank, chron.shape[0], chron.shape[1], offset=rank * per_rank)
    rho = torch.from_numpy(_solve(gm, chron, scen, t0, step=1))
    gather_list = [torch.empty_like(rho) for _ in range(world)] if rank == 0 else None
    dist.gather(rho, gather_list, dst=0)          # the one collective of the path (results to the agent's rank)
    if rank == 0:
        q.put(torch.cat(gather_list).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank():
    world, per_rank = 2, 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gm = GridModel.from_npz(os.path.join(GOLD, "gridmodel_l2rpn_case14_sandbox.npz"))
    chron = np.load(os.path.join(GOLD, "case14_sandbox_chronics.npz"))["chron"]
    scen, t0 = instance_schedule(world * per_rank, chron.shape[0], chron.shape[1])
    want = _solve(gm, chron, scen, t0, step=1)
    assert got.shape == want.shape
    assert np.array_equal(got, want)              # sharding must not change a single bit


def _ring_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from grid2op_b200.collect import RingCollector
    K, n_steps, shape = 4, 23, (5, 3)
    col = RingCollector(rank, world, K, shape, "cpu")
    for k in range(n_steps):
        col.slot(k).copy_(torch.full(shape, float(1000 * rank + k)))      # what the step's kernel would store
        col.step_done(k)
    col.drain()
    if rank == 0:
        ok = True
        for r in range(world):
            for k in range(n_steps - K, n_steps):           # the last K steps are always available on the agent's rank
                ok = ok and bool((col.gathered(r, k) == float(1000 * r + k)).all())
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_ring_collector_two_ranks():
    """RingCollector (bench.py's N > 1 result collection: device ring, one asynchronous gather per K steps, partial gather at the
    end) with gloo on CPU tensors"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ring_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    assert q.get(timeout=180) is True
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0


def test_schedule_is_a_partition():
    s_all, t_all = instance_schedule(4096 * 8, 3, 576)
    for r in range(8):
        s, t = instance_schedule(4096, 3, 576, offset=r * 4096)
        assert np.array_equal(s, s_all[r * 4096:(r + 1) * 4096]) and np.array_equal(t, t_all[r * 4096:(r + 1) * 4096])
