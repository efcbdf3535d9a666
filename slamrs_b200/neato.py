"""Neato XV-11 lidar recordings -> Observations: the real-scan producer in front of
GridMapSlam::update (host-side mirror of slamrs/neato/src/frame.rs).

    NeatoFrame            frame.rs:7-12   (distance mm, strength, valid per degree)
    parse_packets         frame.rs:136-208 (0xFA-framed 22-byte packets, checksum, revolutions)
    load_neato_binary     frame.rs:210-217
    NeatoFrame.observation  `impl From<NeatoFrame> for Observation`, frame.rs:219-238

The packet scan is sequential by nature (a byte that fails the checksum shifts the framing by one),
so it stays a host loop over candidate offsets; field extraction of the accepted packets is
vectorised. ~2 KB per revolution: nothing here belongs on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from .slam import Observation

PACKET = 22


@dataclass
class NeatoFrame:
    distance: np.ndarray   # u16[360], millimetres
    strength: np.ndarray   # u16[360]
    valid: np.ndarray      # u8[360]

    def observation(self, id: int = 0) -> Observation:
        n = self.distance.size
        angle = np.deg2rad(np.arange(n, dtype=np.float64))          # (i as f64).to_radians()
        return Observation(id, angle=angle, distance=self.distance.astype(np.float64) / 1000.0,
                           valid=self.valid != 0)


def _checksum_ok(pkt: np.ndarray) -> bool:
    words = pkt[0:20:2].astype(np.uint32) | (pkt[1:20:2].astype(np.uint32) << 8)
    chk = 0
    for w in words.tolist():
        chk = ((chk << 1) + w) & 0xFFFFFFFF
    chk = ((chk & 0x7FFF) + (chk >> 15)) & 0x7FFF
    return chk == (int(pkt[20]) | (int(pkt[21]) << 8))


def parse_packets(buf: bytes) -> List[NeatoFrame]:
    """All completed revolutions of a recording (the trailing partial one is dropped, as in the reference)."""
    b = np.frombuffer(bytes(buf), np.uint8)
    frames: List[NeatoFrame] = []
    starts: List[int] = []      # offsets of the accepted packets of the current revolution
    slots: List[int] = []       # their packet index 0..89
    last_index = 0

    def flush():
        dist = np.zeros(360, np.uint16); stren = np.zeros(360, np.uint16); valid = np.zeros(360, np.uint8)
        if starts:
            st = np.asarray(starts)[:, None]
            # later packets with the same index overwrite earlier ones, like the array store in the reference
            order = np.arange(len(slots))
            sl = np.asarray(slots)
            last = {s: k for k, s in zip(order.tolist(), sl.tolist())}
            keep = np.array(sorted(last.values()), np.int64)
            st, sl = st[keep], sl[keep]
            off = st + 4 + 4 * np.arange(4)[None, :]                       # first byte of each of the 4 readings
            b0, b1, b2, b3 = b[off], b[off + 1], b[off + 2], b[off + 3]
            cell = (sl[:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
            dist[cell] = (b0.astype(np.uint16) | ((b1.astype(np.uint16) & 0x3F) << 8)).reshape(-1)
            stren[cell] = ((b3.astype(np.uint16) << 8) | b2.astype(np.uint16)).reshape(-1)
            valid[cell] = ((b1 & 0x80) == 0).astype(np.uint8).reshape(-1)
        frames.append(NeatoFrame(dist, stren, valid))

    candidates = np.nonzero(b[: max(0, b.size - PACKET + 1)] == 0xFA)[0]
    for i in candidates.tolist():
        pkt = b[i:i + PACKET]
        if not _checksum_ok(pkt):
            continue
        idx = int(pkt[1])
        if idx < 0xA0:
            continue
        index = idx - 0xA0
        if index >= 90:
            raise ValueError(f"packet index {idx:#x} out of range (the reference would panic)")
        if index < last_index:
            flush()
            starts, slots = [], []
        starts.append(i); slots.append(index)
        last_index = index
    return frames


def load_neato_binary(path: str) -> List[NeatoFrame]:
    with open(path, "rb") as f:
        return parse_packets(f.read())
