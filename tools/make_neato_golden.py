#!/usr/bin/env python
"""Generates tests/golden/neato_out2_head.bin + neato_out2_head.npz from the reference's own Neato
recording (slamrs/baseui/data/out2.bin): the first 24 KiB of the byte stream and the revolutions
the oracle restatement of slamrs/neato/src/frame.rs decodes from them. Run in the build container
(the recording is not available on the GPU box); the outputs are committed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import neato_oracle as NO  # noqa: E402

SRC = "/root/reference/slamrs/baseui/data/out2.bin"
HEAD = 24 * 1024


def main():
    buf = open(SRC, "rb").read()[:HEAD]
    frames = NO.parse_packets(buf)
    out = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(out, "neato_out2_head.bin"), "wb") as f:
        f.write(buf)
    np.savez_compressed(os.path.join(out, "neato_out2_head.npz"),
                        distance=np.array([fr["distance"] for fr in frames], np.uint16),
                        strength=np.array([fr["strength"] for fr in frames], np.uint16),
                        valid=np.array([fr["valid"] for fr in frames], np.uint8))
    print(len(buf), "bytes ->", len(frames), "revolutions; valid per revolution:",
          [int(sum(fr["valid"])) for fr in frames])


if __name__ == "__main__":
    main()
