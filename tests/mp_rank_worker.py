"""One rank of a multi-PROCESS sharded filter (tests/test_gpu_multi_process.py): creates its shard on
GPU `rank`, runs the scans and writes what it saw to an .npz. Ranks in separate processes map each
other's pools through CUDA IPC (slamrs_b200/csrc/api.cu setup_peers), the path torchrun benchmarks use."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def at_scale(rank, world, n, steps, flags, nccl_id, out_path):
    """Back-to-back steps (step_async) of a large sharded filter; writes the final index vector, poses and a sample of
    grids (tests/test_gpu_multi_process.py::test_processes_at_scale...)."""
    from slamrs_b200 import GpuPlacement, GridMapSlam
    from common import SEED, at_scale_config, at_scale_scans
    with GridMapSlam(at_scale_config(n), GpuPlacement(device=rank, rank=rank, world_size=world, nccl_id=nccl_id, seed=SEED,
                                                     flags=flags)) as g:
        for obs, odo in at_scale_scans(steps):
            g.upload_scan(obs)
            g.step_async(odo)
        g.sync()
        rec = {"idx": g.resample_indices(), "poses": g.poses(), "pulled": np.array([int(g.step_history(0, steps)[:, 1].sum())])}
        for p in range(g.first, g.first + g.n_local, 13):
            rec[f"cells_{p}"] = g.cells(p)
    np.savez(out_path, **rec)


def main():
    rank, world, n, steps, flags = (int(a) for a in sys.argv[1:6])
    nccl_id = bytes.fromhex(sys.argv[6])
    out_path = sys.argv[7]
    if len(sys.argv) > 8 and sys.argv[8] == "at_scale":
        return at_scale(rank, world, n, steps, flags, nccl_id, out_path)
    from slamrs_b200 import GpuPlacement, GridMapSlam, GridMapSlamConfig
    from common import SEED, make_scans
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=n)
    scans = make_scans(1.0, 360, 1.0, steps)
    probe = sorted({0, 1, n // 2 - 1, n // 2, n - 1})
    rec = {}
    with GridMapSlam(cfg, GpuPlacement(device=rank, rank=rank, world_size=world, nccl_id=nccl_id, seed=SEED, flags=flags)) as g:
        for step, (obs, odo) in enumerate(scans):
            g.update(obs, odo)
            ep = g.estimated_pose()
            rec[f"idx{step}"] = g.resample_indices()
            rec[f"poses{step}"] = g.poses()
            rec[f"maxp{step}"] = np.array([g.max_particle])
            rec[f"est{step}"] = np.array([ep.x, ep.y, ep.theta], np.float32)
            rec[f"map{step}"] = g.estimated_likelihood().data
            rec[f"pulled{step}"] = np.array([g.stats()["grids_pulled"]])
            for p in probe:
                if g.first <= p < g.first + g.n_local:
                    rec[f"cells{step}_{p}"] = g.cells(p)
    np.savez(out_path, **rec)


if __name__ == "__main__":
    main()
