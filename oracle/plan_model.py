"""TEST INFRASTRUCTURE ONLY -- numpy model of the resampling planner's contract (k_plan in
slamrs_b200/csrc/kernels.cu): how one rank turns the replicated, non-decreasing index vector of
systematic resampling (slamrs/slam/src/grid/particle.rs:78-105) into "keep in place" / "copy"
decisions and a new slot table, without ever writing a slot that a peer GPU copies from in the same
step. Slot assignment is fully determined by ordered lists, so the device's slot tables must equal
this model's, step by step (tests/test_gpu_plan_model.py); the invariants are checked on the CPU
for random index vectors (tests/test_plan_model.py). Only tests/ may import this module."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class PlanResult:
    slot_new: np.ndarray        # S: physical slot of every new local particle
    spare_new: np.ndarray       # E: free slots carried to the next step
    classes: np.ndarray         # S: 0 keep, 1 local copy, 2 remote first use, 3 remote further use
    copies: List[tuple]         # (m, source particle, destination slot), in output order
    leaders: List[int]          # positions in `copies` that start a fan-out sub-run (<= 16 destinations)
    unsafe_slots: np.ndarray    # dropped local slots that another rank copies from in this step
    staging_short: int


def plan(idx: np.ndarray, rank: int, world: int, slot_old: np.ndarray, spare: np.ndarray, fan: int = 16) -> PlanResult:
    idx = np.asarray(idx, np.int64)
    n = idx.size
    s = n // world
    lo, hi = rank * s, (rank + 1) * s
    idx_l = idx[lo:hi]
    slot_old = np.asarray(slot_old, np.int64)
    spare = np.asarray(spare, np.int64)
    first = np.ones(s, bool)
    first[1:] = idx_l[1:] != idx_l[:-1]                       # first use HERE (m == 0 counts as first)
    local = (idx_l >= lo) & (idx_l < hi)
    classes = np.where(local & first, 0, np.where(local, 1, np.where(first, 2, 3)))
    slot_new = np.full(s, -1, np.int64)
    kept = np.zeros(s, bool)
    for m in np.nonzero(classes == 0)[0]:
        kept[idx_l[m] - lo] = True
        slot_new[m] = slot_old[idx_l[m] - lo]
    # a dropped slot is unsafe when some index OUTSIDE this rank's output range selects its particle
    outside = np.concatenate([idx[:lo], idx[hi:]])
    selected_outside = np.zeros(s, bool)
    sel = outside[(outside >= lo) & (outside < hi)] - lo
    selected_outside[sel] = True
    safe = [int(slot_old[j]) for j in range(s) if not kept[j] and not selected_outside[j]]
    unsafe = [int(slot_old[j]) for j in range(s) if not kept[j] and selected_outside[j]]
    free_list = safe + [int(x) for x in spare] + unsafe
    usable = len(safe) + len(spare)
    copies, leaders = [], []
    pos = 0
    run_first = 0
    short = 0
    for m in range(s):
        if m > 0 and idx_l[m] != idx_l[m - 1]:
            run_first = m
        if classes[m] == 0:
            continue
        if pos < usable:
            slot_new[m] = free_list[pos]
            k = (m - run_first - 1) if classes[m] == 1 else (m - run_first)
            if k % fan == 0:
                leaders.append(len(copies))
            copies.append((m, int(idx_l[m]), int(free_list[pos])))
        else:
            short += 1
        pos += 1
    used = min(pos, usable)
    left = usable - used
    e = len(spare)
    spare_new = np.array([free_list[used + i] if i < left else free_list[usable + (i - left)] for i in range(e)], np.int64)
    return PlanResult(slot_new, spare_new, classes, copies, leaders, np.array(unsafe, np.int64), short)


@dataclass
class DeferredResult:
    plan: PlanResult
    materialized: List[tuple]   # (local particle j, root slot, own slot), in particle order: copies before the ray update
    mat_leaders: List[int]      # positions in `materialized` that start a fan-out sub-run
    pulls: List[tuple]          # (m, source particle, destination slot): first local uses of remote sources (real copies)
    alias: np.ndarray           # alias table after the step


def plan_deferred(idx: np.ndarray, rank: int, world: int, slot_old: np.ndarray, spare: np.ndarray, alias: np.ndarray,
                  all_particles: bool = False, fan: int = 16) -> DeferredResult:
    """Deferred copies (PlanArgs::alias_of, k_materialize_list): slot tables are those of plan(); a clone is an
    alias of its source's slot until its particle is about to be written."""
    idx = np.asarray(idx, np.int64)
    s = idx.size // world
    lo, hi = rank * s, (rank + 1) * s
    slot_old = np.asarray(slot_old, np.int64)
    alias = np.array(alias, np.int64)
    selected = np.ones(s, bool) if all_particles else np.isin(np.arange(lo, hi), idx)
    mat, leaders = [], []
    run_start = 0
    for j in range(s):
        own = int(slot_old[j])
        if selected[j] and alias[own] != own:
            if not mat or mat[-1][1] != int(alias[own]):
                run_start = len(mat)
            if (len(mat) - run_start) % fan == 0:
                leaders.append(len(mat))
            mat.append((j, int(alias[own]), own))
    for _, _, own in mat:
        alias[own] = own
    p = plan(idx, rank, world, slot_old, spare, fan)
    pulls = []
    first_copy = {}
    for m, src, dslot in p.copies:
        cls = p.classes[m]
        if cls == 1:
            alias[dslot] = slot_old[src - lo]
        elif cls == 2:
            alias[dslot] = dslot
            first_copy[src] = dslot
            pulls.append((m, src, dslot))
        else:
            alias[dslot] = first_copy[src]
    return DeferredResult(p, mat, leaders, pulls, alias)


def systematic_indices(weights: np.ndarray, u01: float) -> np.ndarray:
    """particle.rs:78-101 on normalised weights (sequential form)."""
    n = weights.size
    r = u01 * 1.0 / n
    c = weights[0]
    i = 0
    out = np.zeros(n, np.int64)
    for m in range(1, n + 1):
        u = r + (m - 1.0) * 1.0 / n
        while u > c and i < n - 1:
            i += 1
            c += weights[i]
        out[m - 1] = i
    return out
