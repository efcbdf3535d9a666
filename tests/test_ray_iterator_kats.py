"""Hand-derived known answers for GridRayIterator (slamrs/slam/src/grid/ray.rs:21-110) and
inverse_sensor_model (map.rs:148-172). The reference has no test for either; these cases were
worked out by hand from the Rust source and pin both restatements (C oracle and numpy)."""
import numpy as np
import pytest

from oracle import numpy_restatement as NP


def _both(oracle, *args, **kw):
    a = [tuple(int(v) for v in c) for c in oracle.ray_cells(*args, **kw)]
    b = NP.ray_cells(*args, **kw)
    assert a == b
    return a


def test_horizontal_ray_with_two_extra_cells(oracle):
    # dy == 0 -> error = e - inf = -inf -> always step x; n = 1 + 2 + (floor(5.5) - 2) = 6
    assert _both(oracle, 2.5, 3.5, 5.5, 3.5, 10, 10) == [(2, 3), (3, 3), (4, 3), (5, 3), (6, 3), (7, 3)]


def test_vertical_ray_downwards(oracle):
    # dx == 0 -> error = +inf -> always step y (y_inc = -1); n = 3 + (6 - 4) = 5
    assert _both(oracle, 1.5, 6.5, 1.5, 4.5, 10, 10) == [(1, 6), (1, 5), (1, 4), (1, 3), (1, 2)]


def test_zero_length_ray_emits_start_cell_three_times(oracle):
    # dx == dy == 0 -> error = inf - inf = NaN -> x-branch with x_inc = 0; n = 1 + extra = 3
    assert _both(oracle, 4.25, 7.75, 4.25, 7.75, 10, 10) == [(4, 7)] * 3
    assert _both(oracle, 4.25, 7.75, 4.25, 7.75, 10, 10, extra=0) == [(4, 7)]


def test_exact_diagonal(oracle):
    # (0.5,0.5)->(3.5,3.5): delta=(3,3), error = 0.5*3 - 0.5*3 = 0 -> first step is x (error > 0 false)
    want = [(0, 0), (1, 0), (1, 1), (2, 1), (2, 2), (3, 2), (3, 3), (4, 3), (4, 4)]
    assert _both(oracle, 0.5, 0.5, 3.5, 3.5, 10, 10) == want


def test_start_outside_grid_yields_nothing(oracle):
    assert _both(oracle, -0.5, 2.5, 5.5, 2.5, 10, 10) == []
    assert _both(oracle, 10.0, 2.5, 5.5, 2.5, 10, 10) == []
    assert _both(oracle, 2.5, 12.0, 2.5, 5.0, 10, 10) == []


def test_ray_stops_permanently_when_it_leaves_the_grid(oracle):
    # heading right out of a 4-wide grid: cells 2,3 then x=4 is outside -> stop (n would allow more)
    assert _both(oracle, 2.5, 1.5, 9.5, 1.5, 4, 4) == [(2, 1), (3, 1)]


def test_shallow_slope(oracle):
    # (0.5,0.5)->(4.5,1.5): delta=(4,1); error = 0.5*1 - 0.5*4 = -1.5
    # steps: x(e=-0.5) x(e=0.5) y(e=-3.5) x x x x ... n = 3 + 4 + 1 = 8
    want = [(0, 0), (1, 0), (2, 0), (2, 1), (3, 1), (4, 1), (5, 1), (6, 1)]
    assert _both(oracle, 0.5, 0.5, 4.5, 1.5, 10, 10) == want


def test_negative_direction(oracle):
    # (4.5,4.5)->(1.5,3.5): delta=(3,1); x_inc=-1: error = (4.5-4)*1 = 0.5; y_inc=-1: error -= (4.5-4)*3 -> -1.0
    # n = 3 + (4-1) + (4-3) = 7
    want = [(4, 4), (3, 4), (2, 4), (2, 3), (1, 3), (0, 3)]
    assert _both(oracle, 4.5, 4.5, 1.5, 3.5, 10, 10) == want  # 7th cell would be x=-1 -> outside, stop


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_rays_both_restatements_agree(oracle, seed):
    rng = np.random.default_rng(seed)
    for _ in range(300):
        w = int(rng.integers(4, 64))
        x0, y0 = rng.uniform(-1, w + 1, 2)
        x1, y1 = rng.uniform(-w, 2 * w, 2)
        if rng.random() < 0.2:
            x1 = x0
        if rng.random() < 0.2:
            y1 = y0
        if rng.random() < 0.2:
            x0 = float(np.floor(x0))
        _both(oracle, np.float32(x0), np.float32(y0), np.float32(x1), np.float32(y1), w, w)


def test_inverse_sensor_model_bands(oracle):
    ism = oracle.inverse_sensor_model  # 0 prior, 1 free, 2 occupied
    # miss: free strictly before the range reading, prior from there on
    assert ism(3.0, 10.0, False) == 1 and ism(10.0, 10.0, False) == 0 and ism(12.0, 10.0, False) == 0
    # hit, tolerance 2.0 -> band [md - 1, md + 1] is occupied, inclusive at both ends
    assert ism(8.99, 10.0, True) == 1
    assert ism(9.0, 10.0, True) == 2 and ism(10.0, 10.0, True) == 2 and ism(11.0, 10.0, True) == 2
    assert ism(11.01, 10.0, True) == 0
    # NaN distance compares false everywhere -> occupied for a hit, prior for a miss
    assert ism(float("nan"), 10.0, True) == 2 and ism(float("nan"), 10.0, False) == 0
