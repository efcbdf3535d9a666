"""The C oracle against the independent numpy restatement (written from SURVEY.md Appendix A):
bit-identical poses, weights, indices and grids over several lockstep updates, small cases."""
import numpy as np
import pytest

from oracle import numpy_restatement as NP

from common import SEED, make_scans


@pytest.mark.parametrize("n,width,res,scale,rng_range,steps", [
    (5, 1.28, 0.04, 0.5, 0.5, 3),     # 32x32 grid, robot inside a small room
    (3, 0.8, 0.05, 1.0, 1.0, 2),      # 16x16 grid smaller than the room: rays leave the grid
])
def test_c_oracle_equals_numpy_restatement(oracle, n, width, res, scale, rng_range, steps):
    pos = (-width / 2, -width / 2)
    scans = make_scans(scale, 72, rng_range, steps)
    c = oracle.OracleSlam(pos, width, width, res, n, True)
    p = NP.Slam(pos, width, width, res, n)
    assert (c.gw, c.gh) == (p.w, p.h)
    for step, (obs, odo) in enumerate(scans):
        z = oracle.motion_normals(SEED, step, 0, n)
        u = oracle.resample_uniform(SEED, step)
        ang = obs.angle.astype(np.float32).astype(np.float64)
        dist = obs.distance.astype(np.float32).astype(np.float64)
        dl, dr, wb = np.float32(odo.distance_left), np.float32(odo.distance_right), np.float32(odo.wheel_distance)
        assert c.update(ang, dist, obs.valid.astype(np.uint8), dl, dr, wb, z, u) == 0
        p.update(ang, dist, obs.valid, dl, dr, wb, z, u)
        w, raw = c.weights()
        assert np.array_equal(raw, np.array(p.raw)), step
        assert np.array_equal(w, np.array(p.weights)), step
        assert list(c.indices()) == p.idx
        assert c.max_particle == p.max_particle
        poses = c.poses()
        for i in range(n):
            assert np.array_equal(poses[i], np.array(p.pose[i], np.float32))
            assert np.array_equal(c.odds(i), p.grid[i]), (step, i)
        # counters reproduce the f64 log-odds (the device representation), to rounding
        for i in range(n):
            nf, no = c.counts(i)
            recon = nf.astype(np.float64) * NP.L_FREE + no.astype(np.float64) * NP.L_OCC
            assert np.allclose(recon, c.odds(i), rtol=1e-13, atol=1e-13)
    c.close()
