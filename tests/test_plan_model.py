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


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("concentration", [0.05, 0.5, 5.0])
@pytest.mark.parametrize("all_particles", [False, True], ids=["survivors-only", "strict-order"])
def test_deferred_copies_show_every_particle_the_grid_eager_copies_would(world, concentration, all_particles):
    """Deferred copies (clones alias their source's slot until written) simulated on symbolic grid contents,
    all ranks together, in the device's order of work: materialise -> integrate -> pull. Before every step each
    particle must see exactly the content that `new[m] = clone(old[i_m])` semantics give it, no copy may write
    a slot whose content is still needed, and roots are always private."""
    rng = np.random.default_rng(7 + 100 * world + int(concentration * 10) + (1000 if all_particles else 0))
    n, s = 48 * world, 48
    e = s if world > 1 else 0
    slot_of = [np.arange(s) for _ in range(world)]
    spare = [np.arange(s, s + e) for _ in range(world)]
    alias = [np.arange(s + e) for _ in range(world)]
    content = [{k: ("prior",) for k in range(s + e)} for _ in range(world)]
    ref = [("prior",)] * n                                   # eager semantics: one private grid per particle
    total_mat = total_eager = 0
    for step in range(14):
        for r in range(world):                               # what the likelihood kernel reads
            for j in range(s):
                root = alias[r][slot_of[r][j]]
                assert alias[r][root] == root, "a root must be private"
                assert content[r][int(root)] == ref[r * s + j]
        idx = _random_indices(rng, n, concentration)
        res = [PM.plan_deferred(idx, r, world, slot_of[r], spare[r], alias[r], all_particles) for r in range(world)]
        selected = set(range(n)) if all_particles else set(idx.tolist())
        # 1. materialise (before anything is integrated on that rank): reads roots, writes the particles' own slots
        for r, d in enumerate(res):
            written = set()
            for j, root, own in d.materialized:
                assert (r * s + j) in selected and root != own and own not in written
                assert root not in written, "a copy reads a slot another copy of the same launch writes"
                content[r][own] = content[r][root]
                written.add(own)
            for a, b in zip(d.mat_leaders, d.mat_leaders[1:] + [len(d.materialized)]):
                assert 1 <= b - a <= 16 and len({d.materialized[k][1] for k in range(a, b)}) == 1
            # the planner hands out slots concurrently: none of them is a slot this step's survivors own,
            # except (strict order only) slots of particles that are integrated and then dropped
            handed = {c[2] for c in d.plan.copies}
            if not all_particles:
                owned = {int(slot_of[r][j]) for j in range(s) if (r * s + j) in selected}
                assert not (handed & owned)
            total_mat += len(d.materialized) + len(d.pulls)
            total_eager += len(d.plan.copies)
        # 2. integrate the scan into the selected particles' own (now private) slots
        upd = {}
        for r in range(world):
            for j in range(s):
                g = r * s + j
                if g in selected:
                    own = int(slot_of[r][j])
                    assert res[r].alias[own] == own or own in {c[2] for c in res[r].plan.copies}
                    content[r][own] = ("scan", step, g, content[r][own])
                    upd[g] = content[r][own]
        # 3. pulls after the barrier: the peer's slot of the source particle, post-update
        for r, d in enumerate(res):
            for m, src, dslot in d.pulls:
                owner = src // s
                assert owner != r
                content[r][dslot] = content[owner][int(slot_of[owner][src - owner * s])]
                assert dslot not in res[r].plan.unsafe_slots.tolist()
        ref = [upd[int(i)] for i in idx]
        slot_of = [d.plan.slot_new for d in res]
        spare = [d.plan.spare_new for d in res]
        alias = [d.alias for d in res]
    if not all_particles:
        assert total_mat < total_eager                       # the point of deferring
