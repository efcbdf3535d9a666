"""CPU model of the device's exact left fold (oracle/fold_model.py) against the plain sequential fold:
the design must be exact on every input and must not need the single-thread fallback on the weight
distributions a particle filter produces (the GPU test test_gpu_resample_exact.py checks the kernel itself)."""
import numpy as np
import pytest

from oracle.fold_model import exact_fold, sequential_fold


def _norm(w):
    return w / sequential_fold(w)[-1]


def _same(a, b):
    return np.array_equal(np.asarray(a).view(np.int64), np.asarray(b).view(np.int64))


@pytest.mark.parametrize("n", [1, 2, 3, 7, 1000, 8192, 20000])
def test_model_is_exact_and_needs_no_fallback(n):
    rng = np.random.default_rng(n)
    cases = {
        "uniform": (rng.random(n), False),
        "equal_normalised": (np.full(n, 1.0) / float(n), True),
        "lognormal20_raw": (np.exp(rng.normal(0, 20, n)), False),
        "tiny_raw": (np.exp(rng.normal(-400, 30, n)), False),
        "all_zero": (np.zeros(n), False),
    }
    cases["lognormal5_normalised"] = (_norm(np.exp(rng.normal(0, 5, n))), True)
    peaked = np.exp(rng.normal(-38, 1, n)); peaked[int(rng.integers(0, max(1, n // 4)))] = 1.0
    cases["one_dominant_particle_normalised"] = (_norm(peaked), True)   # the running sum creeps along 1.0
    for name, (v, fia) in cases.items():
        out, info = exact_fold(v, 8192, fia)
        assert _same(out, sequential_fold(v, fia)), name
        assert not info["fallback"], (name, info)
        assert info["rounds"] <= 2, (name, info)
        assert info["heads"] <= 64, (name, info)


def test_model_falls_back_but_stays_exact_on_hostile_input():
    rng = np.random.default_rng(5)
    w = rng.random(8192); w[100] = np.nan
    out, info = exact_fold(w, 8192)
    assert info["fallback"] and _same(out, sequential_fold(w))
    w = rng.random(8192); w[77] = np.inf
    out, info = exact_fold(w, 8192)
    assert info["fallback"] and _same(out, sequential_fold(w))
    w = 2.0 ** np.arange(0, 300).astype(float)          # every element changes the binade
    out, info = exact_fold(w, 8192)
    assert info["fallback"] and _same(out, sequential_fold(w))
