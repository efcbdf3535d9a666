"""Resample indices bit-exact by construction (particle.rs:40-56, 78-101): the device's k_weights +
k_resample_indices on caller-supplied raw weights against the oracle's strict left folds -- normalised
weights, running sum and indices must agree bit for bit, at the populations of BASELINE.json's configs[3]
and configs[4] and on weights whose prefix sums sit exactly on the resampling thresholds."""
import numpy as np
import pytest

from slamrs_b200.slam import debug_resample

pytestmark = pytest.mark.gpu

U_MAX = 1.0 - 2.0 ** -53      # largest value rand::random::<f64>() returns


def _check(oracle, w, u, expect_no_fallback=True):
    ref = oracle.resample_fold(w, u)
    got = debug_resample(w, u)
    assert np.array_equal(ref["norm"].view(np.int64), got["norm"].view(np.int64)), "normalised weights differ"
    assert np.array_equal(ref["cum"].view(np.int64), got["cum"].view(np.int64)), \
        f"running sum differs first at {np.nonzero(ref['cum'].view(np.int64) != got['cum'].view(np.int64))[0][:4]}"
    assert np.array_equal(ref["idx"].astype(np.int64), got["idx"].astype(np.int64))
    assert ref["max_particle"] == got["max_particle"]
    assert bool(ref["clamped"]) == bool(got["clamped"])
    if expect_no_fallback:
        assert got["fold_fallback"] == 0, got
    return got


@pytest.mark.parametrize("n", [1, 2, 3, 7, 1000, 8192, 65536, 262144])
def test_indices_bit_exact_for_realistic_weights(oracle, n):
    rng = np.random.default_rng(n)
    rounds = []
    for u in (0.0, 0.3718, U_MAX):
        rounds.append(_check(oracle, rng.random(n), u)["fold_rounds"])
        for sigma in (2.0, 10.0):
            rounds.append(_check(oracle, np.exp(rng.normal(-200.0, sigma, n)), u)["fold_rounds"])
    assert max(rounds) <= 3


@pytest.mark.parametrize("n", [8, 4096, 65536, 262144])
def test_prefix_sums_on_the_thresholds(oracle, n):
    """Equal weights: the running sum is k/N up to rounding and with U = 0 every threshold is k/N up to
    rounding: every comparison `u > c` is decided in the last bit."""
    for u in (0.0, 2.0 ** -53, 0.5, U_MAX):
        _check(oracle, np.ones(n), u)
        _check(oracle, np.full(n, 3.0), u)
    # thresholds engineered onto prefixes of uneven weights: w_i = number of thresholds it should take
    rng = np.random.default_rng(n)
    w = rng.integers(0, 4, n).astype(np.float64)
    _check(oracle, w, 0.0)
    _check(oracle, w, U_MAX)


@pytest.mark.parametrize("n", [8192, 65536])
def test_one_dominant_particle(oracle, n):
    """A peaked filter: the running sum reaches 1 - O(ulp) early and creeps along the binade edge 1.0."""
    rng = np.random.default_rng(7 * n)
    for trial in range(6):
        w = np.exp(rng.normal(-38.0, 1.0 + trial, n))
        w[int(rng.integers(0, n // 4))] = 1.0
        got = _check(oracle, w, float(rng.random()))
        assert got["fold_rounds"] <= 3


def test_degenerate_weights_follow_the_reference(oracle):
    """particle.rs:49-56, 91: a zero or non-finite sum makes every weight NaN, `u > c` is then always false
    and every new particle is a copy of particle 0."""
    n = 8192
    got = _check(oracle, np.zeros(n), 0.25, expect_no_fallback=False)   # every weight underflowed: 0/0 = NaN weights
    assert np.all(got["idx"] == 0) and np.all(np.isnan(got["norm"]))
    rng = np.random.default_rng(3)
    w = rng.random(n); w[1234] = np.nan
    got = _check(oracle, w, 0.25, expect_no_fallback=False)
    assert np.all(got["idx"] == 0)
    w = rng.random(n); w[4321] = np.inf                            # finite / inf = 0, inf / inf = NaN
    got = _check(oracle, w, 0.25, expect_no_fallback=False)
    assert got["idx"].max() <= 4321
    w = np.zeros(n); w[n - 1] = 1e-300                             # a single survivor, the last particle
    got = _check(oracle, w, 0.9)
    assert np.all(got["idx"] == n - 1)
    w = np.zeros(n); w[17] = 5.0; w[18] = 5.0                      # an exact tie: the last maximum wins (particle.rs:40-46)
    got = _check(oracle, w, 0.5)
    assert got["max_particle"] == 18 and set(np.unique(got["idx"])) == {17, 18}


def test_index_past_the_end_is_clamped_and_reported(oracle):
    """particle.rs:91-93: rounding can leave c < u after the last weight; the reference then indexes out of
    bounds (panic). The device clamps to N-1 and reports it."""
    hits = 0
    for n in (6, 7, 10, 13, 14, 15, 19, 22):
        ref = oracle.resample_fold(np.ones(n), U_MAX)
        got = _check(oracle, np.ones(n), U_MAX)
        hits += ref["clamped"]
        assert got["idx"][-1] == n - 1
    assert hits > 0, "no case exercised the clamp"
    rng = np.random.default_rng(0)
    for _ in range(20):
        _check(oracle, rng.random(int(rng.integers(2, 5000))), U_MAX)


def test_hostile_input_takes_the_sequential_fallback_and_stays_exact(oracle):
    w = 2.0 ** np.arange(0, 900).astype(np.float64)       # every addition changes the binade
    got = _check(oracle, w, 0.5, expect_no_fallback=False)
    assert got["fold_fallback"] != 0
