"""GPU tests of the rows either side of the step: scan production on the device,
real Neato scans through the filter, the cheaper map read-outs and the effective particle count."""
import os

import numpy as np
import pytest

from slamrs_b200 import GpuPlacement, GridMapSlam, GridMapSlamConfig, Odometry
from slamrs_b200 import _lib
from slamrs_b200 import neato as N
from slamrs_b200.simulator import reference_scene

from common import lockstep, make_scans, oracle_slam, oracle_step

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("scale,n_beams,rng_m,pose", [
    (1.0, 360, 1.0, (0.0, 0.0, 0.0)), (1.0, 360, 1.0, (0.31, -0.52, 2.4)), (10.0, 720, 6.0, (3.2, -7.7, -1.1)),
    (5.0, 90, 100.0, (0.0, 0.0, 0.5)), (1.0, 360, 1.0, (5.0, 5.0, 0.3)),   # outside the room: many rays hit nothing
])
def test_device_lidar_is_bit_exact(oracle, scale, n_beams, rng_m, pose):
    seg = reference_scene(scale)
    ang, dist, valid = oracle.sim_scan(seg, pose, n_beams, rng_m)
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=2)
    with GridMapSlam(cfg) as g:
        n = g.sim_scan(seg, pose, n_beams, rng_m)
        obs = g.get_scan()
    assert n == len(ang) == len(obs)
    assert np.array_equal(obs.angle.astype(np.float32).view(np.uint32), np.asarray(ang, np.float32).view(np.uint32))
    assert np.array_equal(obs.distance.astype(np.float32).view(np.uint32), np.asarray(dist, np.float32).view(np.uint32))
    assert np.array_equal(obs.valid, np.asarray(valid, bool))


def test_step_on_device_produced_scan_equals_step_on_host_scan(oracle):
    """sim_scan + step_async (no host scan at all) must give the same filter as update(host scan)."""
    seg = reference_scene(1.0)
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=16)
    odo = Odometry.new(0.08, 0.10, 0.1)
    poses = [(0.0, 0.0, 0.0), (0.09, 0.01, 0.2), (0.17, 0.04, 0.4)]
    with GridMapSlam(cfg) as a, GridMapSlam(cfg) as b:
        for p in poses:
            a.sim_scan(seg, p, 360, 1.0)
            a.step_async(odo); a.sync()
            ang, dist, valid = oracle.sim_scan(seg, p, 360, 1.0)
            from slamrs_b200 import Observation
            b.update(Observation(0, angle=ang, distance=dist, valid=valid), odo)
            assert np.array_equal(a.poses().view(np.uint32), b.poses().view(np.uint32))
            assert np.array_equal(a.resample_indices(), b.resample_indices())
        for q in (0, 7, 15):
            assert np.array_equal(a.cells(q), b.cells(q))


def test_real_neato_scans_parity(oracle):
    """Twelve revolutions of the reference's own Neato recording through the filter, lockstep with
    the oracle: 5 m x 5 m map at 2.5 cm, 24 particles."""
    frames = N.parse_packets(open(os.path.join(GOLD, "neato_out2_head.bin"), "rb").read())
    assert len(frames) == 12
    odo = Odometry.new(0.004, 0.006, 0.1)
    scans = [(fr.observation(k), odo) for k, fr in enumerate(frames[:6])]
    cfg = GridMapSlamConfig(position=(-6.4, -6.4), width=12.8, height=12.8, resolution=0.05, n_particles=24)
    errs = lockstep(oracle, cfg, scans, particles=range(0, 24, 5))
    print(errs[-1])


def test_map_window_formats_and_extent(oracle):
    cfg = GridMapSlamConfig(position=(-12.8, -12.8), width=25.6, height=25.6, resolution=0.05, n_particles=12)
    scans = make_scans(5.0, 360, 6.0, 3)
    with GridMapSlam(cfg) as g:
        assert g.map_extent() == (0, 0, 0, 0)
        for obs, odo in scans:
            g.update(obs, odo)
        full = g.estimated_likelihood().data.reshape(512, 512)
        x0, y0, x1, y1 = g.map_extent()
        assert 0 <= x0 < x1 <= 512 and 0 <= y0 < y1 <= 512 and x0 % 8 == 0 and x1 % 8 == 0
        outside = np.ones((512, 512), bool); outside[y0:y1, x0:x1] = False
        assert np.all(full[outside] == 0.5)                      # nothing informed outside the extent
        assert (x1 - x0) * (y1 - y0) < 512 * 512 // 2            # and the extent is much smaller than the grid
        win, w64 = g.estimated_likelihood_window(fmt=_lib.MAP_F64)
        assert win == (x0, y0, x1, y1) and np.array_equal(w64, full[y0:y1, x0:x1])
        _, w32 = g.estimated_likelihood_window(fmt=_lib.MAP_F32)
        assert np.array_equal(w32, full[y0:y1, x0:x1].astype(np.float32))
        _, w8 = g.estimated_likelihood_window(fmt=_lib.MAP_U8)
        assert np.array_equal(w8, np.rint(full[y0:y1, x0:x1] * 255.0).astype(np.uint8))
        _, part = g.estimated_likelihood_window(window=(8, 16, 40, 17), fmt=_lib.MAP_F32)
        assert np.array_equal(part, full[16:17, 8:40].astype(np.float32))
        with pytest.raises(_lib.SlamrsGpuError):
            g.estimated_likelihood_window(window=(0, 0, 520, 8))


def test_number_of_effective_particles(oracle):
    cfg = GridMapSlamConfig(position=(-2.0, -2.0), width=4.0, height=4.0, resolution=0.02, n_particles=64)
    scans = make_scans(1.0, 360, 1.0, 4)
    osl = oracle_slam(oracle, cfg)
    with GridMapSlam(cfg) as g:
        assert g.number_of_effective_particles() == 64.0
        for step, (obs, odo) in enumerate(scans):
            oracle_step(oracle, osl, obs, odo, step)
            g.update(obs, odo)
            ref = osl.number_of_effective_particles()
            got = g.number_of_effective_particles()
            assert abs(got - ref) <= 1e-9 * ref, (step, got, ref)
            assert 1.0 <= got <= 64.0
    osl.close()


def test_pipelined_map_readout_equals_the_blocking_one():
    """estimated_likelihood_async(t) + update(t+1) + map_wait(): the map of step t, bit for bit, although the next
    step was issued (and has changed the estimate's grid) before the copy was awaited; two read-outs in flight."""
    import torch
    cfg = GridMapSlamConfig(position=(-4.0, -4.0), width=8.0, height=8.0, resolution=0.02, n_particles=64)
    scans = make_scans(2.0, 360, 2.0, 7)
    n = 400 * 400
    bufs = [torch.empty(n, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
    with GridMapSlam(cfg) as a, GridMapSlam(cfg) as b:
        assert a.grid_w * a.grid_h == n
        want = []
        for obs, odo in scans:
            b.update(obs, odo)
            want.append(b.estimated_likelihood().data.copy())
        for t, (obs, odo) in enumerate(scans):
            a.update(obs, odo)
            if t >= 2:      # the buffer about to be reused still holds map t-2 (two read-outs pending at most)
                a.map_wait()
                assert np.array_equal(bufs[t & 1], want[t - 2])
                assert np.array_equal(bufs[(t - 1) & 1], want[t - 1])
            bufs[t & 1][:] = -1.0
            a.estimated_likelihood_async(bufs[t & 1])
        a.map_wait()
        assert np.array_equal(bufs[(len(scans) - 1) & 1], want[-1])
        assert np.array_equal(a.estimated_likelihood().data, want[-1])
