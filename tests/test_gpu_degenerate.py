"""The reference's quirks at the edges of the filter, in full steps on the device:
every weight underflowing to zero (NaN weights, every new particle a copy of particle 0), a single
survivor, and a stationary robot whose start cell saturates its 16-bit counter."""
import numpy as np
import pytest

from slamrs_b200 import GpuPlacement, GridMapSlam, GridMapSlamConfig, Observation, Odometry
from slamrs_b200 import _lib
from slamrs_b200.slam import L_FREE, L_OCC

from common import SEED, compare_step, make_scans, oracle_slam

pytestmark = pytest.mark.gpu


def _step_both(oracle, gpu, osl, obs, odo, z, u):
    ang = obs.angle.astype(np.float32).astype(np.float64)
    dist = obs.distance.astype(np.float32).astype(np.float64)
    rc = osl.update(ang, dist, obs.valid.astype(np.uint8), np.float32(odo.distance_left), np.float32(odo.distance_right),
                    np.float32(odo.wheel_distance), z, u)
    gpu.update(obs, odo, z_draws=z, resample_u=u)
    return rc


def test_all_weights_underflow_then_recovery(oracle):
    """particle.rs:49-56, 91: motion draws 60 sigma out make every weight exp(-inf) = 0, the sum is 0, every
    normalised weight 0/0 = NaN, `u > c` is never true and every new particle is a copy of old particle 0.
    The next (ordinary) step must continue from that generation in lockstep."""
    n = 64
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=n)
    scans = make_scans(1.0, 360, 1.0, 4)
    osl = oracle_slam(oracle, cfg)
    rng = np.random.default_rng(5)
    with GridMapSlam(cfg, GpuPlacement(seed=SEED, rng_mode=_lib.RNG_CALLER)) as gpu:
        for step, (obs, odo) in enumerate(scans):
            z = rng.standard_normal(2 * n)
            if step == 1:
                z[0::2] = 60.0                      # centre-distance draw: pdf underflows to 0, log -> -inf
            rc = _step_both(oracle, gpu, osl, obs, odo, z, float(rng.random()))
            assert rc == 0
            if step == 1:
                w_ref, raw_ref = osl.weights()
                w, raw = gpu.weights()
                assert np.all(raw_ref == 0.0) and np.all(raw == 0.0)
                assert np.all(np.isnan(w_ref)) and np.all(np.isnan(w))
                assert np.all(osl.indices() == 0) and np.all(gpu.resample_indices() == 0)
                assert gpu.max_particle == osl.max_particle
                assert np.array_equal(osl.poses().view(np.uint32), gpu.poses().view(np.uint32))
                for p in (0, 1, n - 1):
                    nf, no = osl.counts(p); gf, go = gpu.counts(p)
                    assert np.array_equal(nf, gf) and np.array_equal(no, go)
            else:
                compare_step(gpu, osl)
    osl.close()


def test_single_survivor(oracle):
    """One particle keeps a sane draw, every other one is thrown 40 sigma away: a single source for the
    whole new generation (N-1 clones of one grid), then an ordinary step."""
    n = 48
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=n)
    scans = make_scans(1.0, 360, 1.0, 4)
    osl = oracle_slam(oracle, cfg)
    rng = np.random.default_rng(6)
    with GridMapSlam(cfg, GpuPlacement(seed=SEED, rng_mode=_lib.RNG_CALLER)) as gpu:
        for step, (obs, odo) in enumerate(scans):
            z = rng.standard_normal(2 * n)
            if step == 2:
                z[0::2] = 38.0
                z[2 * 17] = 0.1                     # particle 17 survives
            rc = _step_both(oracle, gpu, osl, obs, odo, z, float(rng.random()))
            assert rc == 0
            compare_step(gpu, osl)
            if step == 2:
                assert np.all(gpu.resample_indices() == 17)
    osl.close()


def test_stationary_robot_saturates_the_start_cell():
    """A robot that does not move casts 360 rays from the same start cell every scan: the cell takes 360
    free updates per scan and its 16-bit counter saturates at 65,535 in scan 183 (182 * 360 = 65,520).
    Until then the reconstructed log-odds is n * ln(0.3/0.7) as in the reference's f64 sum; from then on
    the counter stays at 65,535 -- the probability is 0 either way (exp underflows long before), only the
    log-odds value stops following the reference. The step reports it (stats.counter_saturated)."""
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=2)
    obs, _ = make_scans(1.0, 360, 1.0, 1)[0]
    still = Odometry(0.0, 0.0, 0.1)
    z = np.zeros(4)                                  # no motion noise: the start cell never changes
    first_saturated = None
    with GridMapSlam(cfg, GpuPlacement(seed=SEED, rng_mode=_lib.RNG_CALLER)) as gpu:
        sx = int((0.0 - cfg.position[0]) / cfg.resolution)
        cell = sx * gpu.grid_h + sx
        for scan in range(1, 200):
            gpu.update(obs, still, z_draws=z, resample_u=0.5)
            if gpu.stats()["counter_saturated"] and first_saturated is None:
                first_saturated = scan
            if scan in (1, 100, 182, 183, 199):
                nf, no = gpu.counts(0)
                assert int(nf[cell]) == min(65535, 360 * scan) and int(no[cell]) == 0
                lo = gpu.log_odds(0)[cell]
                assert abs(lo - min(65535, 360 * scan) * L_FREE) <= 1e-9 * abs(lo)
                assert gpu.estimated_likelihood().data[cell] == 0.0 if scan > 3 else True
    assert first_saturated == 183
