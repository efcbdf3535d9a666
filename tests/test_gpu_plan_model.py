"""The device planner (k_plan) against its numpy model (oracle/plan_model.py): slot tables and spare
lists must be identical after every step, on one GPU and sharded over two (ranks as threads)."""
import threading

import numpy as np
import pytest

from oracle import plan_model as PM
from slamrs_b200 import GpuPlacement, GridMapSlam, GridMapSlamConfig, nccl_unique_id

from common import SEED, make_scans

pytestmark = pytest.mark.gpu


EAGER = 16   # SLAMRS_FLAG_EAGER_COPY


@pytest.mark.parametrize("flags", [0, EAGER, 2, 2 | EAGER], ids=["deferred", "eager", "deferred-strict", "eager-strict"])
def test_slot_tables_match_model_single_gpu(flags):
    cfg = GridMapSlamConfig(position=(-1.28, -1.28), width=2.56, height=2.56, resolution=0.04, n_particles=2048)
    scans = make_scans(1.0, 360, 1.0, 6)
    with GridMapSlam(cfg, GpuPlacement(flags=flags)) as g:
        slot_old, spare = g.slots()
        alias = np.arange(2048)
        assert np.array_equal(slot_old, np.arange(2048)) and spare.size == 0
        for obs, odo in scans:
            g.update(obs, odo)
            if flags & EAGER:
                p = PM.plan(g.resample_indices(), 0, 1, slot_old, spare)
                n_copied, n_leaders = len(p.copies), len(p.leaders)
            else:   # the slot tables are the same; only the clones that are about to be written are copied
                d = PM.plan_deferred(g.resample_indices(), 0, 1, slot_old, spare, alias, all_particles=bool(flags & 2))
                p, alias = d.plan, d.alias
                # strict order: ordered list with fan-out sub-runs; survivors only: every copy reads its own source
                n_copied = len(d.materialized) + len(d.pulls)
                n_leaders = (len(d.mat_leaders) if flags & 2 else len(d.materialized)) + len(d.pulls)
            slot_new, spare_new = g.slots()
            assert np.array_equal(slot_new, p.slot_new)
            st = g.stats()
            assert st["grids_copied"] == n_copied and st["distinct_sources"] == int((p.classes == 0).sum())
            hist = g.step_history(st["step"] - 1, 1)
            assert hist[0, 3] == n_leaders                           # source reads = fan-out sub-runs
            slot_old = slot_new


@pytest.mark.parametrize("flags", [0, EAGER], ids=["deferred", "eager"])
def test_slot_tables_match_model_two_gpus(flags):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, world = 256, 2
    cfg = GridMapSlamConfig(position=(-1.28, -1.28), width=2.56, height=2.56, resolution=0.04, n_particles=n)
    scans = make_scans(1.0, 360, 1.0, 6)
    nid = nccl_unique_id()
    rec = [None] * world
    errs = []

    def worker(rank):
        g = None
        try:
            g = GridMapSlam(cfg, GpuPlacement(device=rank, rank=rank, world_size=world, nccl_id=nid, seed=SEED, spare_slots=n // world,
                                              flags=flags))
            out = [(None,) + g.slots()]
            for obs, odo in scans:
                g.update(obs, odo)
                out.append((g.resample_indices().copy(),) + g.slots() + (g.stats(),))
            rec[rank] = out
        except Exception as e:  # noqa: BLE001
            errs.append((rank, repr(e)))
        finally:
            if g is not None:
                g.close()

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert not errs, errs
    remote = 0
    for r in range(world):
        slot_old, spare = rec[r][0][1], rec[r][0][2]
        alias = np.arange(n // world + n // world)
        for step in range(1, len(scans) + 1):
            idx, slot_new, spare_new, st = rec[r][step]
            if flags & EAGER:
                p = PM.plan(idx, r, world, slot_old, spare)
                n_copied = len(p.copies)
            else:
                d = PM.plan_deferred(idx, r, world, slot_old, spare, alias)
                p, alias = d.plan, d.alias
                n_copied = len(d.materialized) + len(d.pulls)
            assert p.staging_short == 0
            assert np.array_equal(slot_new, p.slot_new), (r, step)
            assert np.array_equal(spare_new, p.spare_new), (r, step)
            assert st["grids_copied"] == n_copied
            assert st["grids_pulled"] == int((p.classes == 2).sum())
            remote += int((p.classes >= 2).sum())
            slot_old, spare = slot_new, spare_new
    assert remote > 0
