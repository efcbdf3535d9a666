"""Invariants of the resampling planner's contract on random index vectors, world sizes 1, 2, 4
(CPU; the model is compared with the device's slot tables in tests/test_gpu_plan_model.py)."""
import numpy as np
import pytest

from oracle import plan_model as PM


def _random_indices(rng, n, concentration):
    w = rng.gamma(concentration, size=n)
    w /= w.sum()
    return PM.systematic_indices(w, float(rng.uniform()))


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("concentration", [0.05, 0.5, 5.0])
def test_planner_invariants(world, concentration):
    rng = np.random.default_rng(100 * world + int(concentration * 10))
    n, s = 64 * world, 64
    e = s if world > 1 else 0
    slot_old = [np.arange(s) for _ in range(world)]
    spare = [np.arange(s, s + e) for _ in range(world)]
    for step in range(12):
        idx = _random_indices(rng, n, concentration)
        assert np.all(np.diff(idx) >= 0)
        plans = [PM.plan(idx, r, world, slot_old[r], spare[r]) for r in range(world)]
        for r, p in enumerate(plans):
            assert p.staging_short == 0                                   # full spare: never short
            all_slots = np.concatenate([p.slot_new, p.spare_new])
            assert np.array_equal(np.sort(all_slots), np.arange(s + e))    # every slot exactly once
            dst = np.array([c[2] for c in p.copies], np.int64)
            # nothing written in this step is a slot a peer copies from in this step, nor a kept slot
            assert not np.intersect1d(dst, p.unsafe_slots).size
            kept_slots = p.slot_new[p.classes == 0]
            assert not np.intersect1d(dst, kept_slots).size
            assert len(dst) == len(set(dst.tolist()))
            # conservation: consumers = dropped local slots (particle.rs:97-100 writes N new particles)
            assert len(p.copies) == s - (p.classes == 0).sum()
            # fan-out sub-runs: every copy belongs to a leader's run of <= 16 copies of one source
            for a, b in zip(p.leaders, p.leaders[1:] + [len(p.copies)]):
                assert 1 <= b - a <= 16 and len({p.copies[k][1] for k in range(a, b)}) == 1
            # the unsafe slots are exactly the dropped local particles selected by another rank
            lo = r * s
            outside = np.concatenate([idx[:lo], idx[lo + s:]])
            want = sorted(int(slot_old[r][j]) for j in range(s)
                          if (lo + j) in set(outside.tolist()) and (lo + j) not in set(idx[lo:lo + s].tolist()))
            assert sorted(p.unsafe_slots.tolist()) == want
        slot_old = [p.slot_new for p in plans]
        spare = [p.spare_new for p in plans]


def test_staging_short_is_reported_not_hidden():
    """One spare slot, two dropped slots that the peer copies from: one consumer finds no writable slot."""
    s = 8
    idx = np.array([0, 0, 0, 0, 0, 1, 2, 3] + [3, 3, 3, 3, 3, 3, 3, 15], np.int64)   # rank 1 pulls particle 3 from rank 0
    p0 = PM.plan(idx, 0, 2, np.arange(s), np.array([8]))
    p1 = PM.plan(idx, 1, 2, np.arange(s), np.array([8]))
    assert p0.staging_short == 0 and p0.unsafe_slots.size == 0            # particle 3 is kept on rank 0
    assert p1.staging_short == 0
    idx2 = np.array([0] * 8 + [5, 5, 6, 6, 7, 7, 15, 15], np.int64)       # rank 0 drops 5, 6, 7; rank 1 copies from them
    q0 = PM.plan(idx2, 0, 2, np.arange(s), np.array([8]))
    assert sorted(q0.unsafe_slots.tolist()) == [5, 6, 7]
    assert q0.staging_short == 2                                          # 7 consumers, 4 safe + 1 spare usable
