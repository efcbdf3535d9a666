"""Pins of the oracle against what the reference itself holds for the path:
the three unit tests of slamrs/common/src/math.rs:159-195, the 222-valid-beam scan recorded in
slamrs/out.log:4, and published Philox4x32-10 known answers for the shared stream."""
import math

import numpy as np
import pytest


def test_math_rs_inverse_roundtrip(oracle):
    # math.rs:167-177 `inverse`: Probability -> LogOdds -> Probability, epsilon 1e-6, v/100 for v in 0..100
    for v in range(100):
        p = v / 100.0
        with np.errstate(all="ignore"):
            back = oracle.log_odds_probability(oracle.prob_log_odds(p))
        assert back == pytest.approx(p, abs=1e-6)


def test_math_rs_zero_is_half(oracle):
    # math.rs:180-182 `zero_is_half`
    assert oracle.prob_log_odds(0.5) == 0.0


def test_math_rs_angle_diff(oracle):
    # math.rs:185-194 `test_angle_diff`, the 8 cases verbatim
    PI = math.pi
    cases = [(PI, PI, 0.0), (-PI, PI, 0.0), (0.0, PI, -PI), (PI, 0.0, -PI), (0.0, PI / 2, PI / 2),
             (PI / 2, 0.0, -PI / 2), (PI, PI / 2, -PI / 2), (PI / 2, PI, PI / 2)]
    for a, b, want in cases:
        assert oracle.angle_diff(a, b) == pytest.approx(want, rel=1e-12, abs=1e-12)


def test_inverse_sensor_model_constants(oracle):
    # map.rs:154-156 through Probability::log_odds (math.rs:30-32)
    assert oracle.prob_log_odds(0.30) == -0.8472978603872036
    assert oracle.prob_log_odds(0.9) == 2.1972245773362196
    assert oracle.prob_log_odds(0.5) == 0.0


def test_first_simulator_scan_has_222_valid_beams(oracle):
    # slamrs/out.log:4 records 222 points for the first scan at the origin of the shipped scene
    from slamrs_b200.simulator import reference_scene
    a, d, v = oracle.sim_scan(reference_scene(1.0), (0.0, 0.0, 0.0), 360, 1.0)
    assert len(a) == 360 and int(v.sum()) == 222


def test_philox4x32_10_known_answers(oracle):
    # Random123 kat_vectors: philox4x32 10 rounds
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        assert [int(x) for x in oracle.philox(ctr, key)] == want


def test_shared_stream_statistics_and_accuracy(oracle):
    z = oracle.motion_normals(0x5EED5A11, 3, 0, 200_000)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs(np.mean(z[0::2] * z[1::2])) < 0.01
    assert abs(np.mean(z ** 4) - 3.0) < 0.1
    for x in [1.0, 0.5, 0.1, 1e-5, 2.0 ** -53, 0.7071, 0.99999]:
        assert oracle.dlog(x) == pytest.approx(math.log(x), rel=2e-15, abs=1e-300)
    for u in [0.0, 0.1, 0.25, 0.3, 0.5, 0.77, 0.999999]:
        s, c = oracle.dsincos2pi(u)
        assert s == pytest.approx(math.sin(2 * math.pi * u), abs=5e-16)
        assert c == pytest.approx(math.cos(2 * math.pi * u), abs=5e-16)
    us = [oracle.resample_uniform(7, s) for s in range(2000)]
    assert 0.0 <= min(us) and max(us) < 1.0 and abs(np.mean(us) - 0.5) < 0.03
