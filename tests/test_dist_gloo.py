"""world_size-2 gloo test of the multi-GPU host logic: shard the population, run each shard's
cells (here with the C oracle standing in for the per-GPU kernel), reduce the summaries with the
same torch.distributed code bench.py uses, and check against the unsharded run."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, N, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ekf4fastcharge_b200"))
    import torch.distributed as dist
    from mek import bank as Bk, dist as D, population as P
    from mek._lib import STAT_NAMES
    from oracle import coracle
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    bank = Bk.Bank.load(Bk.default_bank_path())
    lo, hi = D.shard_range(B, rank, world)
    pop = P.make_population(bank.params[2], hi - lo, seed=7, first_cell=lo)
    co = coracle.COracle(bank)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    lg = co.run(0, N)
    local = {k: 0.0 for k in STAT_NAMES}
    local.update(cells=float(hi - lo), cellsteps=float((hi - lo) * N), sum_nexec=float(lg["nexec"].sum()),
                 n_qp=float((lg["nexec"] > 0).sum()), n_cap=float((lg["nexec"] >= 100).sum()),
                 min_phise=float(lg["phise"].min()), max_V=float(lg["V"].max()), sum_soc=float(lg["SOC"][-1].sum()))
    tot = D.reduce_summaries(local)
    tmax = D.max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((tot, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_summary_reduce(bank, coracle_mod):
    from mek import population as P
    B, N, world = 10, 40, 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    tot, tmax = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    pop = P.make_population(bank.params[2], B, seed=7)
    co = coracle_mod.COracle(bank)
    co.init(pop.SOC0, pop.T, pop.SigmaX0, pop.limits, pop.ekf_SOC0)
    lg = co.run(0, N)
    assert tot["cells"] == B and tot["cellsteps"] == B * N
    assert tot["sum_nexec"] == lg["nexec"].sum() and tot["n_qp"] == (lg["nexec"] > 0).sum()
    assert tot["min_phise"] == lg["phise"].min() and tot["max_V"] == lg["V"].max()
    assert abs(tot["sum_soc"] - lg["SOC"][-1].sum()) < 1e-12
    assert tmax == 11.0
