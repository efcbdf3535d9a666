#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (the Rust reference cannot run here and has
no golden vectors of its own for the grid module -- parity unpinned, see oracle/slam_oracle.c).
The fixtures freeze the oracle's behaviour so that drift in either the oracle or the CUDA path
is caught:  python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O            # noqa: E402
from common import SEED, make_scans       # noqa: E402

CASES = {
    # name: (n_particles, width, resolution, scene scale, beams, range, steps)
    "preset_10x360_200": (10, 4.0, 0.02, 1.0, 360, 1.0, 6),
    "small_32x90_64": (32, 2.56, 0.04, 1.0, 90, 1.0, 8),
    "room_16x360_256_range6": (16, 12.8, 0.05, 5.0, 360, 6.0, 4),
}


def run(name):
    n, width, res, scale, beams, rng, steps = CASES[name]
    pos = (-width / 2, -width / 2)
    scans = make_scans(scale, beams, rng, steps)
    osl = O.OracleSlam(pos, width, width, res, n, True)
    out = {"meta": np.array([n, beams, steps, osl.gw], np.int64), "cfg": np.array([pos[0], pos[1], width, res], np.float64),
           "seed": np.array([SEED], np.uint64)}
    for s, (obs, odo) in enumerate(scans):
        z = O.motion_normals(SEED, s, 0, n)
        u = O.resample_uniform(SEED, s)
        ang = obs.angle.astype(np.float32); dist = obs.distance.astype(np.float32)
        osl.update(ang.astype(np.float64), dist.astype(np.float64), obs.valid.astype(np.uint8),
                   np.float32(odo.distance_left), np.float32(odo.distance_right), np.float32(odo.wheel_distance), z, u)
        w, raw = osl.weights()
        out[f"s{s}_angle"] = ang; out[f"s{s}_dist"] = dist; out[f"s{s}_valid"] = obs.valid.astype(np.uint8)
        out[f"s{s}_odo"] = np.array([odo.distance_left, odo.distance_right, odo.wheel_distance], np.float32)
        out[f"s{s}_z"] = z; out[f"s{s}_u"] = np.array([u])
        out[f"s{s}_poses"] = osl.poses(); out[f"s{s}_raw"] = raw; out[f"s{s}_w"] = w
        out[f"s{s}_idx"] = osl.indices().astype(np.uint32); out[f"s{s}_max"] = np.array([osl.max_particle], np.int64)
        out[f"s{s}_est_pose"] = osl.estimated_pose()
        # grids: exact counters of the estimate's particle + a position-weighted checksum of every grid
        nf, no = osl.counts(osl.max_particle)
        nz = np.nonzero(nf.astype(np.uint32) | no.astype(np.uint32))[0].astype(np.uint32)
        out[f"s{s}_est_cells_idx"] = nz
        out[f"s{s}_est_cells_val"] = (nf[nz].astype(np.uint32) | (no[nz].astype(np.uint32) << 16))
        sums = []
        for p in range(n):
            a, b = osl.counts(p)
            k = np.arange(a.size, dtype=np.uint64) + np.uint64(1)
            sums.append([int((a.astype(np.uint64) * k).sum() % (1 << 61)), int((b.astype(np.uint64) * k).sum() % (1 << 61))])
        out[f"s{s}_grid_checksums"] = np.array(sums, np.uint64)
    osl.close()
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    for c in CASES:
        run(c)
