"""Filter extras the reference leaves as ideas, behind settings that default to the
reference's behaviour: adaptive resampling on the effective number of particles (particle.rs:59-65 computes
it, nothing calls it) and a uniform, global-localisation-style start (README.md:45). Both have oracle
counterparts (oracle/slam_oracle.c so_set_adaptive_resampling, oracle/shared_stream.c ss_uniform_pose) and
run in lockstep with them."""
import numpy as np
import pytest

from slamrs_b200 import GpuPlacement, GridMapSlam, GridMapSlamConfig

from common import SEED, WEIGHT_RTOL, compare_step, make_scans, oracle_slam, oracle_step

pytestmark = pytest.mark.gpu


def test_adaptive_resampling_lockstep(oracle):
    n, tau = 48, 0.5
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=n)
    scans = make_scans(1.0, 360, 1.0, 10)
    osl = oracle_slam(oracle, cfg)
    osl.set_adaptive_resampling(tau)
    decisions = []
    with GridMapSlam(cfg, GpuPlacement(seed=SEED, resample_threshold=tau)) as gpu:
        for step, (obs, odo) in enumerate(scans):
            rc, _, _ = oracle_step(oracle, osl, obs, odo, step)
            assert rc == 0
            gpu.update(obs, odo)
            st = gpu.stats()
            assert bool(st["resampled"]) == osl.resampled(), (step, gpu.number_of_effective_particles(), osl.number_of_effective_particles())
            decisions.append(osl.resampled())
            if not osl.resampled():
                assert np.array_equal(gpu.resample_indices(), np.arange(n, dtype=np.uint32))
                assert st["particles_integrated"] == n          # nobody is dropped: every grid receives the scan
            compare_step(gpu, osl)
            assert abs(gpu.number_of_effective_particles() - osl.number_of_effective_particles()) <= 1e-9 * n
    assert any(decisions) and not all(decisions), decisions
    osl.close()


def test_threshold_zero_is_the_reference(oracle):
    """resample_threshold = 0 (the default) resamples after every update, as slam.rs:74 does."""
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=16)
    with GridMapSlam(cfg, GpuPlacement(seed=SEED)) as gpu:
        for obs, odo in make_scans(1.0, 360, 1.0, 3):
            gpu.update(obs, odo)
            assert gpu.stats()["resampled"] == 1


def test_uniform_init_matches_the_stream_and_runs_in_lockstep(oracle):
    n = 40
    cfg = GridMapSlamConfig(position=(-6.4, -6.4), width=12.8, height=12.8, resolution=0.05, n_particles=n)
    box = (-4.5, -4.0, 4.5, 4.0)
    scans = make_scans(5.0, 360, 3.0, 3)
    osl = oracle_slam(oracle, cfg)
    poses = oracle.uniform_poses(SEED, 0, n, box)
    assert poses[:, 0].min() >= box[0] and poses[:, 0].max() < box[2] and poses[:, 1].min() >= box[1] and poses[:, 1].max() < box[3]
    assert poses[:, 2].min() >= -np.pi - 1e-6 and poses[:, 2].max() <= np.pi
    assert len(np.unique(poses[:, 0])) == n
    osl.set_poses(poses)
    with GridMapSlam(cfg, GpuPlacement(seed=SEED)) as gpu:
        gpu.init_uniform(box)
        assert np.array_equal(gpu.poses().view(np.uint32), poses.view(np.uint32))
        for step, (obs, odo) in enumerate(scans):
            rc, _, _ = oracle_step(oracle, osl, obs, odo, step)
            assert rc == 0
            gpu.update(obs, odo)
            compare_step(gpu, osl)
    osl.close()
