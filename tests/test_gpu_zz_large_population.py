"""More than 8,192 particles per GPU: the planner works in global memory instead of shared memory
(PLAN_STAGED_MAX_S, kernels_resample.cu). Same contract as tests/test_gpu_plan_model.py -- slot tables and
copy counts equal the numpy model after every step -- plus: deferred and eager copies give identical grids.
(Sorted last on purpose: it is the slowest gpu test.)"""
import numpy as np
import pytest

from oracle import plan_model as PM
from slamrs_b200 import GpuPlacement, GridMapSlam, GridMapSlamConfig

from common import make_scans

pytestmark = pytest.mark.gpu

EAGER = 16   # SLAMRS_FLAG_EAGER_COPY
N = 9216


def _run(flags, scans, probe):
    cfg = GridMapSlamConfig(position=(-1.28, -1.28), width=2.56, height=2.56, resolution=0.04, n_particles=N)
    rec = []
    with GridMapSlam(cfg, GpuPlacement(flags=flags)) as g:
        slot_old, spare = g.slots()
        alias = np.arange(N)
        for obs, odo in scans:
            g.update(obs, odo)
            idx = g.resample_indices().copy()
            if flags & EAGER:
                p = PM.plan(idx, 0, 1, slot_old, spare)
                n_copied = len(p.copies)
            else:
                d = PM.plan_deferred(idx, 0, 1, slot_old, spare, alias)
                p, alias = d.plan, d.alias
                n_copied = len(d.materialized)
            slot_new, _ = g.slots()
            assert np.array_equal(slot_new, p.slot_new)
            st = g.stats()
            assert st["grids_copied"] == n_copied and st["distinct_sources"] == int((p.classes == 0).sum())
            slot_old = slot_new
            rec.append((idx, g.poses().copy(), g.weights()[0].copy(), g.estimated_likelihood().data.copy(),
                        [g.cells(q).copy() for q in probe]))
    return rec


def test_global_memory_planner_matches_model_and_modes_agree():
    scans = make_scans(1.0, 360, 1.0, 5)
    probe = (0, 1, 4095, 8191, 8192, N - 1)
    deferred = _run(0, scans, probe)
    eager = _run(EAGER, scans, probe)
    for (i0, p0, w0, m0, c0), (i1, p1, w1, m1, c1) in zip(deferred, eager):
        assert np.array_equal(i0, i1) and np.array_equal(p0.view(np.uint32), p1.view(np.uint32))
        assert np.array_equal(w0.view(np.uint64), w1.view(np.uint64))
        assert np.array_equal(m0, m1)
        for a, b in zip(c0, c1):
            assert np.array_equal(a, b)
        # clones of one source hold identical grids
        dup = np.nonzero(np.diff(i0.astype(np.int64)) == 0)[0]
        assert dup.size > 0
