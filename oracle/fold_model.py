"""TEST INFRASTRUCTURE ONLY -- CPU model of the device's exact left fold (k_weights, kernels_resample.cu).

The reference sums weights strictly left to right (`iter().sum()`, particle.rs:50; `c += weight[i]`,
particle.rs:91-93). A parallel machine cannot re-associate that sum without changing its rounding, and
a resample threshold that lands within a few ulps of a prefix value then selects a different particle.
The device therefore reproduces the sequential fold exactly, in parallel, from two facts:

  * while the running sum s stays inside one binade [2^e, 2^(e+1)], s is an integer multiple S of
    ulp = 2^(e-52) and fl(s + w) - s depends on s only through the parity of S (round-half-even
    ties). A chunk of elements is therefore a two-state transducer (d0, d1): the total increment for
    an even / odd S at entry. Transducers compose associatively, so a prefix scan applies.
  * a chunk in which the sum leaves its binade is a "head": its output is computed by a short
    sequential chain from the exact incoming value.

The binade of every chunk's incoming sum is GUESSED from an ordinary (re-associated) prefix sum; the
result is PROVED by induction: every chunk re-folds its elements from its incoming value with real
additions, and its result must equal, bit for bit, the incoming value its successor derived
independently. Everything in front of the first mismatch is proven; the procedure restarts behind it
from the now exact value (the "anchor"), which also settles on which side of a binade edge a sum that
creeps along the edge lies (normalised weights end within a few ulps of 1.0, a binade edge). After
MAX_ROUNDS rounds, or with too many heads, a plain sequential fold takes over.

This module models the same steps thread by thread (T chunks) so that the design -- in particular
how rarely the sequential fallback is needed -- can be tested without a GPU (tests/test_fold_model.py).
"""
from __future__ import annotations

import math
import struct

import numpy as np

HEADS_PER_CTA = 32     # per 1024-thread CTA (the records travel through distributed shared memory)
CTA_THREADS = 1024
MARGIN_BITS = 40       # "near a binade edge": within this relative distance (the sequential fold may differ
                       # from the re-associated prefix by a few thousand ulps at most for 2^18 elements)
MAX_ROUNDS = 4


def _bits(x: float) -> int:
    return struct.unpack("<q", struct.pack("<d", x))[0]


def _exponent(x: float) -> int:
    return ((_bits(x) >> 52) & 0x7FF) - 1023


def sequential_fold(v, first_is_assignment=False):
    """The reference: s = 0.0; s += v[i] (or s = v[0] first). Returns all prefixes."""
    out = np.empty(len(v), np.float64)
    s = 0.0
    for i, x in enumerate(v):
        s = float(x) if (i == 0 and first_is_assignment) else s + float(x)
        out[i] = s
    return out


class Td:
    __slots__ = ("d0", "d1", "q", "cnt")

    def __init__(self, d0=0.0, d1=0.0, q=0, cnt=0):
        self.d0, self.d1, self.q, self.cnt = d0, d1, q, cnt


def compose(a: Td, b: Td) -> Td:
    """a then b. A range that contains a head forgets everything before its last head."""
    if b.cnt > 0:
        return Td(b.d0, b.d1, b.q, a.cnt + b.cnt)
    p0 = a.q & 1
    d0 = a.d0 + (b.d1 if p0 else b.d0)
    q0 = p0 ^ ((b.q >> p0) & 1)
    a1 = (a.q >> 1) & 1
    p1 = 1 ^ a1
    d1 = a.d1 + (b.d1 if p1 else b.d0)
    q1 = a1 ^ ((b.q >> p1) & 1)
    return Td(d0, d1, q0 | (q1 << 1), a.cnt)


def apply_tail(t: Td, s: float) -> float:
    return s + (t.d1 if (_bits(s) & 1) else t.d0)


def guess_binade(a: float, anchor: float):
    """Binade the running sum is assumed to be in when the re-associated prefix says `a` (> 0).
    Away from the binade edges: a's own. Within the margin of an edge 2^k the sum may be on either
    side: the side of the anchor (the last exactly known value) if the anchor is inside the same
    zone, else the lower side (a sum creeping up to an edge is below it until proven otherwise)."""
    e = _exponent(a)
    x0 = math.ldexp(1.0, e)
    m = math.ldexp(1.0, -MARGIN_BITS)
    if a < x0 * (1.0 + m):
        k = e            # edge 2^e just below a
    elif a > 2.0 * x0 * (1.0 - m):
        k = e + 1        # edge 2^(e+1) just above a
    else:
        return e, False
    edge = math.ldexp(1.0, k)
    if anchor > 0.0 and abs(anchor - edge) <= edge * m:
        return (k if anchor >= edge else k - 1), True
    return k - 1, True


def exact_fold(v, n_threads=8192, first_is_assignment=False):
    """Returns (prefixes, info). info: heads (first round), rounds, fallback (bool), reason."""
    v = np.asarray(v, np.float64)
    n = len(v)
    L = max(1, -(-n // n_threads))
    T = n_threads
    lo = [min(n, t * L) for t in range(T)]
    hi = [min(n, lo[t] + L) for t in range(T)]
    part = [0.0] * T
    zero = [True] * T
    for t in range(T):
        s = 0.0
        for i in range(lo[t], hi[t]):
            s += float(v[i])
            zero[t] = zero[t] and float(v[i]) == 0.0
        part[t] = s
    out = np.empty(n, np.float64)
    info = {"heads": 0, "rounds": 0, "fallback": False, "reason": ""}

    def fold_chunk(t, s):
        for i in range(lo[t], hi[t]):
            s = float(v[i]) if (i == 0 and first_is_assignment) else s + float(v[i])
            out[i] = s
        return s

    t0, s0 = 0, 0.0          # anchor: chunks < t0 are proven, s0 is the exact incoming value of chunk t0
    for rnd in range(MAX_ROUNDS):
        info["rounds"] = rnd + 1
        # ---- re-associated prefix behind the anchor (blocks of 32 summed separately, then combined)
        ain = [0.0] * (T + 1)
        acc = s0
        for w0 in range(0, T, 32):
            wsum = 0.0
            for t in range(w0, min(T, w0 + 32)):
                ain[t] = acc + wsum
                if t >= t0:
                    wsum += part[t]
            acc = acc + wsum
        # ---- per-chunk transducers
        tds = []
        for t in range(T):
            if t < t0 or lo[t] >= hi[t]:
                tds.append(Td())
                continue
            a_in, a_out = ain[t], ain[t] + part[t]
            td = None
            if zero[t]:
                td = Td()                       # adding zeros changes nothing, whatever the binade
            elif a_in > 0.0 and math.isfinite(a_out):
                e, zone_in = guess_binade(a_in, s0)
                e_out, zone_out = guess_binade(a_out, s0)
                # regular: the chunk is assumed to stay inside binade e. A chunk that enters an edge zone
                # from outside is a head (the sum may or may not reach the edge inside it).
                if e == e_out and (zone_in or not zone_out) and -960 <= e <= 1000:
                    x0 = math.ldexp(1.0, e)
                    x1 = x0 + math.ldexp(1.0, e - 52)
                    r0, r1 = x0, x1
                    for i in range(lo[t], hi[t]):
                        r0 += float(v[i]); r1 += float(v[i])
                    if r1 <= 2 * x0:
                        td = Td(r0 - x0, r1 - x1, (_bits(r0) & 1) | (((_bits(r1) & 1) ^ 1) << 1), 0)
            tds.append(td if td is not None else Td(0.0, 0.0, 0, 1))
        excl = []
        run = Td()
        for t in range(T):
            excl.append(run)
            run = compose(run, tds[t])
        heads = [t for t in range(T) if tds[t].cnt]
        if rnd == 0:
            info["heads"] = len(heads)
        per_cta = {}
        for t in heads:
            per_cta[t // CTA_THREADS] = per_cta.get(t // CTA_THREADS, 0) + 1
        if per_cta and max(per_cta.values()) > HEADS_PER_CTA:
            info.update(fallback=True, reason="too many heads")
            break
        # ---- chain over the heads
        s_out_head = []
        for g, t in enumerate(heads):
            base = s_out_head[g - 1] if g > 0 else s0
            s = apply_tail(excl[t], base)
            s_out_head.append(fold_chunk(t, s))
        # ---- every chunk: incoming value, final fold, comparison with the successor's incoming value
        s_in = [0.0] * T
        for t in range(t0, T):
            p = excl[t]
            s_in[t] = apply_tail(p, s_out_head[p.cnt - 1] if p.cnt > 0 else s0)
        assert _bits(s_in[t0]) == _bits(s0)
        last = max([t for t in range(T) if lo[t] < hi[t]], default=-1)
        first_bad = None
        for t in range(t0, last + 1):
            s_out = fold_chunk(t, s_in[t])
            if t < last and _bits(s_out) != _bits(s_in[t + 1]):
                first_bad = (t + 1, s_out)
                break                           # (the device folds every chunk and takes the minimum)
        if first_bad is None:
            return out, info
        t0, s0 = first_bad
        if s0 != s0 or math.isinf(s0):
            info.update(fallback=True, reason="non-finite sum")
            break
    else:
        info.update(fallback=True, reason="rounds exhausted")
    return sequential_fold(v, first_is_assignment), info
