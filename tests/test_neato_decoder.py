"""The Neato lidar decoder (slamrs/neato/src/frame.rs): oracle restatement against the golden
revolutions decoded from the reference's own recording, and the product decoder against the
oracle, including damaged streams. CPU only."""
import os

import numpy as np
import pytest

from oracle import neato_oracle as NO
from slamrs_b200 import neato as N

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden():
    buf = open(os.path.join(GOLD, "neato_out2_head.bin"), "rb").read()
    z = np.load(os.path.join(GOLD, "neato_out2_head.npz"))
    return buf, z


def _same(ref, got):
    assert len(ref) == len(got)
    for r, g in zip(ref, got):
        assert np.array_equal(np.array(r["distance"], np.uint16), g.distance)
        assert np.array_equal(np.array(r["strength"], np.uint16), g.strength)
        assert np.array_equal(np.array(r["valid"], np.uint8), g.valid)


def test_oracle_reproduces_the_golden_revolutions():
    buf, z = _golden()
    frames = NO.parse_packets(buf)
    assert len(frames) == z["distance"].shape[0] == 12
    for k, fr in enumerate(frames):
        assert np.array_equal(np.array(fr["distance"], np.uint16), z["distance"][k])
        assert np.array_equal(np.array(fr["strength"], np.uint16), z["strength"][k])
        assert np.array_equal(np.array(fr["valid"], np.uint8), z["valid"][k])


def test_checksum_known_answer():
    """First packet of the recording, checked by hand: words folded as chk = (chk << 1) + w."""
    buf, _ = _golden()
    i = buf.index(0xFA)
    pkt = buf[i:i + 22]
    words = [pkt[2 * k] | (pkt[2 * k + 1] << 8) for k in range(10)]
    chk = 0
    for w in words:
        chk = (chk << 1) + w
    chk = ((chk & 0x7FFF) + (chk >> 15)) & 0x7FFF
    assert NO.checksum_ok(pkt) == (chk == (pkt[20] | (pkt[21] << 8)))
    bad = bytearray(pkt); bad[5] ^= 0x10
    assert NO.checksum_ok(bytes(pkt)) and not NO.checksum_ok(bytes(bad))


def test_product_decoder_matches_oracle_on_golden_stream():
    buf, _ = _golden()
    _same(NO.parse_packets(buf), N.parse_packets(buf))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_product_decoder_matches_oracle_on_damaged_streams(seed):
    """Dropped bytes, flipped bits, spurious 0xFA bytes and truncation: framing must resynchronise identically."""
    buf, _ = _golden()
    rng = np.random.default_rng(seed)
    b = bytearray(buf[: 12000 + 1000 * seed])
    for _ in range(40):
        k = int(rng.integers(0, len(b)))
        op = int(rng.integers(0, 3))
        if op == 0:
            del b[k]
        elif op == 1:
            b[k] ^= 1 << int(rng.integers(0, 8))
        else:
            b.insert(k, 0xFA)
    _same(NO.parse_packets(bytes(b)), N.parse_packets(bytes(b)))
    _same(NO.parse_packets(b""), N.parse_packets(b""))
    _same(NO.parse_packets(bytes(b[:21])), N.parse_packets(bytes(b[:21])))


def test_observation_conversion():
    buf, _ = _golden()
    fr_ref = NO.parse_packets(buf)[3]
    fr = N.parse_packets(buf)[3]
    ang, dist, _, valid = NO.to_observation(fr_ref)
    obs = fr.observation()
    assert np.array_equal(obs.angle, np.array(ang))          # (i as f64).to_radians(), bit for bit
    assert np.array_equal(obs.distance, np.array(dist))      # mm / 1000.0 in f64
    assert np.array_equal(obs.valid, np.array(valid))
    assert len(obs) == 360
