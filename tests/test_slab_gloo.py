"""CPU, world_size 2 and 3 over gloo: the slab driver (sph-fluid-simulator_b200/slab.py) — cuts,
all-to-all migration, 1-cell ghost halo, density halo — driving an oracle-backed engine, checked
against the oracle on the undivided domain by particle id."""
import importlib
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, steps, rebalance_every, out_dir, fast=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import Oracle
    from slab_cpu_engine import CpuOracleEngine
    slab = importlib.import_module("sph-fluid-simulator_b200.slab")
    g = np.load(os.path.join(ROOT, "tests", "golden", "cube20_step200.npz"))
    O = Oracle()
    s = O.settings(tuple(float(v) for v in g["settings"]))
    pos, vel = g["pos0"], g["vel0"]
    n = pos.shape[0]
    ids = np.arange(n, dtype=np.uint32)
    eng = CpuOracleEngine(O, s)
    # deliberately bad initial distribution: round-robin rows, so the first step migrates almost everything
    mine = ids % world == rank
    eng.upload(pos[mine], vel[mine], ids[mine])
    drv = slab.SlabDriver(eng, rank, world, x_lo=-60, nbins=121)
    drv.rebalance()
    if fast:
        drv.step(float(g["dt"]))  # general path delivers the round-robin rows to their owners
        drv.setup_fast(drv.suggest_halo_rows())
        for k in range(1, steps):
            drv.step_fast(float(g["dt"]))
    else:
        for k in range(steps):
            if rebalance_every and k and k % rebalance_every == 0:
                drv.rebalance()
            drv.step(float(g["dt"]))
    st = eng.owned_state()
    cx = eng.cell_x()[(eng.idw & 0x80000000) == 0]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), cuts=np.array(drv.cuts, np.int64), cell_x=cx,
             migrated=drv.stats["migrated_rows"], halo=drv.stats["halo_rows"], **st)
    dist.destroy_process_group()


def _runner_worker(rank, world, port, steps, out_dir):
    """SlabRunner in steady state over the sync-free path, starting from deliberately lopsided cuts: the
    balance check moves every cut one cell at a time and the sync-free steps absorb each move as migration."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import Oracle
    from slab_cpu_engine import CpuOracleEngine
    slab = importlib.import_module("sph-fluid-simulator_b200.slab")
    g = np.load(os.path.join(ROOT, "tests", "golden", "cube20_step200.npz"))
    O = Oracle()
    s = O.settings(tuple(float(v) for v in g["settings"]))
    pos, vel = g["pos0"], g["vel0"]
    ids = np.arange(pos.shape[0], dtype=np.uint32)
    eng = CpuOracleEngine(O, s)
    mine = ids % world == rank
    eng.upload(pos[mine], vel[mine], ids[mine])
    drv = slab.SlabDriver(eng, rank, world, x_lo=-60, nbins=121)
    run = slab.SlabRunner(drv, float(g["dt"]), transport="nccl", check_every=1, threshold=0.02, halo_slack=4.0)
    # lopsided start: balanced cuts shifted two cells to the left (rank 0 gets too little)
    balanced = drv.rebalance()
    drv.cuts = [balanced[0]] + [c - 2 for c in balanced[1:-1]] + [balanced[-1]]
    drv.rebalance = lambda: drv.cuts  # the runner's first step must keep the lopsided cuts
    first_cuts = list(drv.cuts)
    run.run(steps)
    st = eng.owned_state()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), cuts=np.array(drv.cuts, np.int64), first_cuts=np.array(first_cuts, np.int64),
             balanced=np.array(balanced, np.int64), general=run.general_steps, moves=run.cut_moves, checks=run.balance_checks, **st)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_runner_rebalances_one_cell_at_a_time_without_general_steps(oracle, world):
    steps = 6
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_runner_worker, args=(world, _free_port(), steps, d), nprocs=world, join=True)
        ranks = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]
    want = _oracle_reference(oracle, steps)
    ids = np.concatenate([r["id"] for r in ranks])
    assert np.array_equal(np.sort(ids), np.arange(len(want["pos"]), dtype=np.uint32)), "every particle owned exactly once"
    for d in ranks:
        i = d["id"]
        assert np.array_equal(d["hash"], want["hash"][i])
        assert (np.abs(d["density"] - want["density"][i]) / want["density"][i]).max() < 1e-5
        assert np.abs(d["pos"] - want["pos"][i]).max() < 2e-5
        # exactly one general (all-to-all) step, at step 0; the cuts walked back towards balance one cell per check
        assert int(d["general"]) == 1 and int(d["checks"]) == steps - 1 and int(d["moves"]) >= 2
        inner_first, inner_now, inner_bal = d["first_cuts"][1:-1], d["cuts"][1:-1], d["balanced"][1:-1]
        assert (np.abs(inner_now - inner_bal) < np.abs(inner_first - inner_bal)).all()
        assert (np.abs(inner_now - inner_first) <= steps - 1).all()


def _run(world, steps, rebalance_every=0, fast=False):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), steps, rebalance_every, d, fast), nprocs=world, join=True)
        return [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]


def _oracle_reference(oracle, steps):
    g = load_golden("cube20_step200.npz")
    s = oracle.settings(tuple(float(v) for v in g["settings"]))
    p, v, ids = g["pos0"], g["vel0"], np.arange(g["pos0"].shape[0], dtype=np.uint32)
    for _ in range(steps):
        o = oracle.step(s, float(g["dt"]), p, v, ids)
        p, v, ids = o["pos"], o["vel"], o["id"]
    inv = np.argsort(ids)
    return {k: o[k][inv] for k in ("pos", "vel", "density", "force", "hash")}


@pytest.mark.parametrize("world", [2, 3])
def test_slab_driver_matches_undivided_oracle(oracle, world):
    steps = 3
    ranks = _run(world, steps)
    want = _oracle_reference(oracle, steps)
    ids = np.concatenate([r["id"] for r in ranks])
    assert np.array_equal(np.sort(ids), np.arange(len(want["pos"]), dtype=np.uint32)), "every particle owned exactly once"
    for r, d in enumerate(ranks):
        cuts = d["cuts"]
        i = d["id"]
        assert np.array_equal(d["hash"], want["hash"][i])
        assert (np.abs(d["density"] - want["density"][i]) / want["density"][i]).max() < 1e-5
        assert np.abs(d["pos"] - want["pos"][i]).max() < 1e-5
        assert np.abs(d["vel"] - want["vel"][i]).max() < 1e-4
        fn = np.linalg.norm(want["force"][i], axis=1)
        scale = np.maximum(fn, np.median(fn))
        assert (np.linalg.norm(d["force"] - want["force"][i], axis=1) / scale).max() < 1e-3
        assert len(i) > 0 and d["halo"] > 0
    # the round-robin start forces a large first migration
    assert sum(int(d["migrated"]) for d in ranks) > len(want["pos"]) // 3
    # balanced cuts: no rank holds more than ~1.5x its share
    assert max(len(d["id"]) for d in ranks) < 1.5 * len(want["pos"]) / world + 200


def test_sync_free_step_matches_undivided_oracle(oracle):
    """The fixed-size-message path (step_fast) after one general step, world_size 2."""
    ranks = _run(2, 4, fast=True)
    want = _oracle_reference(oracle, 4)
    ids = np.concatenate([r["id"] for r in ranks])
    assert np.array_equal(np.sort(ids), np.arange(len(want["pos"]), dtype=np.uint32))
    for d in ranks:
        i = d["id"]
        assert np.array_equal(d["hash"], want["hash"][i])
        assert (np.abs(d["density"] - want["density"][i]) / want["density"][i]).max() < 1e-5
        assert np.abs(d["pos"] - want["pos"][i]).max() < 2e-5


def test_rebalance_moves_ownership(oracle):
    ranks = _run(2, 4, rebalance_every=2)
    want = _oracle_reference(oracle, 4)
    ids = np.concatenate([r["id"] for r in ranks])
    assert np.array_equal(np.sort(ids), np.arange(len(want["pos"]), dtype=np.uint32))
    for d in ranks:
        assert np.abs(d["pos"] - want["pos"][d["id"]]).max() < 2e-5


def test_choose_cuts_properties():
    slab = importlib.import_module("sph-fluid-simulator_b200.slab")
    rng = np.random.default_rng(3)
    for world in (1, 2, 4, 8):
        hist = rng.integers(0, 50, 200)
        hist[:20] = 0
        hist[150:] = 0
        cuts = slab.choose_cuts(hist, -100, world)
        assert len(cuts) == world + 1 and cuts[0] == slab.INT_MIN and cuts[-1] == slab.INT_MAX
        inner = cuts[1:-1]
        assert all(b > a for a, b in zip(inner, inner[1:]))
        own = slab.owner_of(cuts, np.arange(-100, 100))
        per = np.bincount(own, weights=hist, minlength=world)
        assert per.max() <= hist.sum() / world + hist.max() + 1
    with pytest.raises(ValueError):
        slab.choose_cuts(np.array([0, 5, 0]), 0, 2)
    # lopsided: everything in two cells, four ranks impossible, two ranks fine
    assert slab.choose_cuts(np.array([0, 9, 9, 0]), 10, 2)[1] == 12
