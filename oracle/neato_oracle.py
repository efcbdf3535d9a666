"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of the reference's Neato XV-11 lidar decoder,
the producer of the real scans that feed GridMapSlam::update.

Follows slamrs/neato/src/frame.rs statement by statement:
    parse_data                      frame.rs:74-83
    calculate_checksum_and_validate frame.rs:85-106
    parse_packet                    frame.rs:108-122
    Revolution::as_readings         frame.rs:43-71
    parse_packets                   frame.rs:136-208
    From<NeatoFrame> for Observation frame.rs:219-238
Pinned by the reference's own recordings (slamrs/baseui/data/out.bin): tests/golden/neato_*.
Only tests/ may import this module.
"""
from __future__ import annotations

import math


def parse_data(b):
    assert len(b) == 4
    return dict(valid=(b[1] & (1 << 7)) == 0, strength_warning=(b[1] & (1 << 6)) == 0,
                distance=b[0] | ((b[1] & 0x3F) << 8), strength=(b[3] << 8) | b[2])


def checksum_ok(b):
    assert len(b) == 22
    words = []
    for i in range((len(b) - 2) // 2):
        words.append((b[2 * i + 1] << 8) | b[2 * i])
    chk32 = 0
    for d in words:
        chk32 = ((chk32 << 1) + d) & 0xFFFFFFFF          # u32 arithmetic (never wraps for 10 words)
    checksum = (chk32 & 0x7FFF) + (chk32 >> 15)
    checksum = checksum & 0x7FFF
    cs = (b[21] << 8) | b[20]
    return checksum == cs


def parse_packet(b):
    assert len(b) == 22
    return dict(index=b[1], speed=(b[3] << 8) | b[2],
                data=[parse_data(b[4:8]), parse_data(b[8:12]), parse_data(b[12:16]), parse_data(b[16:20])],
                checksum=checksum_ok(b))


def as_readings(packets):
    distance = [0] * 360
    strength = [0] * 360
    valid = [0] * 360
    for i, p in enumerate(packets):
        if p is not None:
            for j in range(4):
                distance[i * 4 + j] = p["data"][j]["distance"]
                strength[i * 4 + j] = p["data"][j]["strength"]
                valid[i * 4 + j] = int(p["data"][j]["valid"])
    return dict(distance=distance, strength=strength, valid=valid)


def parse_packets(buf: bytes):
    frames = []
    i = 0
    packets = [None] * 90
    last_index = 0
    while i < len(buf):
        if buf[i] == 0xFA and (len(buf) - i) >= 22:
            p = parse_packet(buf[i:i + 22])
            if not p["checksum"]:
                i += 1
                continue
            if p["index"] < 0xA0:          # checked_sub underflow
                i += 1
                continue
            index = p["index"] - 0xA0
            if index < last_index:
                frames.append(as_readings(packets))
                packets = [None] * 90
            packets[index] = p             # (an index above 89 panics in the reference; recordings never hold one)
            last_index = index
        i += 1
    return frames


def to_observation(frame):
    """-> (angle[360] f64, distance[360] f64, strength[360] f64, valid[360] bool)"""
    ang = [math.radians(float(i)) for i in range(len(frame["distance"]))]   # (i as f64).to_radians()
    dist = [frame["distance"][i] / 1000.0 for i in range(len(frame["distance"]))]
    return ang, dist, [float(s) for s in frame["strength"]], [v != 0 for v in frame["valid"]]
