"""GPU: the slab-decomposed step (CUDA engine + sph-fluid-simulator_b200/slab.py) against the
single-GPU step, by particle id. Cells are ordered by particle id inside the library, so every sum
is taken in the same order whatever the decomposition: the comparison is BIT-EXACT.

With one visible GPU the ranks share device 0 and exchange through gloo with host staging (NCCL
refuses two ranks on one device); with two or more GPUs the same test also runs over NCCL."""
import importlib
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, assert_bit_equal, load_golden

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene(S, jet=False):
    """Dense interacting state with hash-collision double counts: the 20^3 golden cube at step 200.
    jet: a few particles in the middle of the cube cross more than one cell per step along x, which
    the sync-free steps must notice (they otherwise only scan the slab's edge layers)."""
    g = load_golden("cube20_step200.npz")
    s = S.default_settings()
    pos, vel = g["pos0"].copy(), g["vel0"].copy()
    if jet:
        mid = np.argsort(np.abs(pos[:, 0] - np.median(pos[:, 0])))[:24]
        vel[mid[:12], 0] = 1.3 * s.h / s.dt
        vel[mid[12:], 0] = -1.3 * s.h / s.dt
    return s, pos, vel


def _worker(rank, world, port, backend, steps, fast, out_dir, jet=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sph_b200 as S
    slab = importlib.import_module("sph-fluid-simulator_b200.slab")
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    s, pos, vel = _scene(S, jet)
    n = pos.shape[0]
    ids = np.arange(n, dtype=np.uint32)
    drv, sim = slab.make_gpu_driver(s, 2 * n + 65536, dev, rank, world)  # room for the fixed-size message regions
    mine = (ids // 97) % world == rank  # scrambled initial ownership: the first step migrates
    sim.upload(pos[mine], vel[mine], ids[mine])
    drv.rebalance()
    if fast:
        # first step through the general path (it delivers the scrambled rows to their owners),
        # then the sync-free path: fixed-size NCCL/gloo messages, or stores into peer mailboxes
        drv.step(s.dt)
        if fast == "p2p":
            drv.setup_p2p(drv.suggest_halo_rows())
            for k in range(1, steps):
                drv.step_p2p(s.dt)
        else:
            drv.setup_fast(drv.suggest_halo_rows())
            for k in range(1, steps):
                drv.step_fast(s.dt)
    else:
        for k in range(steps):
            if k == steps // 2:
                drv.rebalance()
            drv.step(s.dt)
    sim.sync()
    d = slab.gather_owned(sim)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), cuts=np.array(drv.cuts, np.int64),
             migrated=drv.stats["migrated_rows"], halo=drv.stats["halo_rows"], **d)
    sim.close()
    dist.destroy_process_group()


def _single_gpu_reference(sph, steps, out_dir, jet=False):
    s, pos, vel = _scene(sph, jet)
    sim = sph.Sim(s, capacity=pos.shape[0])
    sim.upload(pos, vel)
    sim.step(steps)
    out = sim.download(sph.ORDER_ID)
    st = sim.stats()
    sim.close()
    assert st.mean_density > 9.6, "reference state should be interacting (selfDens alone is 9.284)"
    return out


def _compare(ranks, want, world):
    ids = np.concatenate([r["id"] for r in ranks])
    assert np.array_equal(np.sort(ids), np.arange(len(want["pos"]), dtype=np.uint32)), "each particle owned once"
    for d in ranks:
        i = d["id"]
        assert len(i) > 0
        assert np.array_equal(d["hash"], want["hash"][i])
        for k in ("pos", "vel", "density", "force"):
            assert_bit_equal(d[k], want[k][i], f"slab vs single GPU: {k}")
    if world > 1:
        assert sum(int(d["migrated"]) for d in ranks) > 0
        assert max(len(d["id"]) for d in ranks) < 1.6 * len(want["pos"]) / world


@pytest.mark.parametrize("fast", [False, True, "p2p"], ids=["general", "syncfree", "peer-mailbox"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_step_is_bit_identical_to_single_gpu_gloo(sph, world, fast):
    steps = 6
    with tempfile.TemporaryDirectory() as d:
        want = _single_gpu_reference(sph, steps, d)
        mp.spawn(_worker, args=(world, _free_port(), "gloo", steps, fast, d), nprocs=world, join=True)
        ranks = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]
    _compare(ranks, want, world)


@pytest.fixture(params=["1", "0"], ids=["edge-scans", "full-scans"])
def edge_scans(request):
    """SPH_B200_EDGE_SCAN for the spawned ranks: 1 (the default) — the sync-free steps look for migrants and
    ghosts in the slab's edge x-layers only; 0 — they look at every row. Same bits either way."""
    old = os.environ.get("SPH_B200_EDGE_SCAN")
    os.environ["SPH_B200_EDGE_SCAN"] = request.param
    yield
    if old is None:
        del os.environ["SPH_B200_EDGE_SCAN"]
    else:
        os.environ["SPH_B200_EDGE_SCAN"] = old


@pytest.mark.parametrize("fast", [True, "p2p"], ids=["syncfree", "peer-mailbox"])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("jet", [False, True], ids=["calm", "jet"])
def test_edge_scans_are_bit_identical_and_fall_back_for_fast_particles(sph, edge_scans, world, fast, jet):
    """jet: a few particles cross more than one cell per step, so an interior row leaves the slab;
    the integration flags it and the next step scans every row."""
    steps = 7
    with tempfile.TemporaryDirectory() as d:
        want = _single_gpu_reference(sph, steps, d, jet=jet)
        mp.spawn(_worker, args=(world, _free_port(), "gloo", steps, fast, d, jet), nprocs=world, join=True)
        ranks = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]
    _compare(ranks, want, world)


def test_overlapped_density_exchange_is_bit_identical(sph):
    """SPH_B200_P2P_OVERLAP=1: the halo densities travel on a second stream while the rows without ghost
    neighbours are integrated, the boundary rows follow in a second launch (opt-in: measured, no gain)."""
    old = os.environ.get("SPH_B200_P2P_OVERLAP")
    os.environ["SPH_B200_P2P_OVERLAP"] = "1"
    try:
        steps, world = 6, 3
        with tempfile.TemporaryDirectory() as d:
            want = _single_gpu_reference(sph, steps, d)
            mp.spawn(_worker, args=(world, _free_port(), "gloo", steps, "p2p", d), nprocs=world, join=True)
            ranks = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]
        _compare(ranks, want, world)
    finally:
        if old is None:
            del os.environ["SPH_B200_P2P_OVERLAP"]
        else:
            os.environ["SPH_B200_P2P_OVERLAP"] = old


@pytest.mark.parametrize("fast", [False, True, "p2p"], ids=["general", "syncfree", "peer-mailbox"])
def test_slab_step_over_nccl(sph, fast):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    steps, world = 6, 2
    with tempfile.TemporaryDirectory() as d:
        want = _single_gpu_reference(sph, steps, d)
        mp.spawn(_worker, args=(world, _free_port(), "nccl", steps, fast, d), nprocs=world, join=True)
        ranks = [dict(np.load(os.path.join(d, f"rank{r}.npz"))) for r in range(world)]
    _compare(ranks, want, world)
