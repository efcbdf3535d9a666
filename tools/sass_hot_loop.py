#!/usr/bin/env python
"""Shape of the ray kernel's free-run loop in the built library's SASS (no GPU needed).

    python tools/sass_hot_loop.py [slamrs_b200/libslamrs_gpu.so] [-v]

The walk of k_ray_update_packed spends its time in one loop: classify the cell, one shared-memory
add, one step of the reference's ray iterator. ptxas compiles it to a single predicated block of 25
instructions -- or, after changes elsewhere in the kernel, to a branchy form that rebuilds the row
table's shared address inside the loop (S2UR + ULEA per y-step, 31 instructions, +15 % kernel time
measured on B200). This script finds the loop around the first ATOMS.ADD of the kernel and prints its
size and whether it contains an S2UR; tests/test_abi_and_host.py fails the build on the slow form."""
import re
import subprocess
import sys

KERNEL = "_ZN6slamrs19k_ray_update_packed"


def hot_loop(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    lines = out.split("\n")
    start = next(i for i, l in enumerate(lines) if "Function : " + KERNEL in l)
    end = next((i for i, l in enumerate(lines[start + 1:], start + 1) if "Function :" in l), len(lines))
    ins = []
    for l in lines[start:end]:
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    k = next(j for j, (_, t) in enumerate(ins) if "ATOMS.ADD" in t)
    for j in range(k, min(k + 120, len(ins))):   # the loop's backward branch
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", ins[j][1])
        if m and int(m.group(1), 16) <= ins[k][0]:
            body = [(a, t) for a, t in ins if int(m.group(1), 16) <= a <= ins[j][0]]
            return {"instructions": len(body), "s2ur": any("S2UR" in t for _, t in body),
                    "local_memory": any(t.startswith(("LDL", "STL")) for _, t in body), "body": body}
    raise RuntimeError("no loop found around the first ATOMS.ADD of " + KERNEL)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-")]
    r = hot_loop(args[0] if args else "slamrs_b200/libslamrs_gpu.so")
    print({k: v for k, v in r.items() if k != "body"})
    if "-v" in sys.argv:
        for a, t in r["body"]:
            print(f"{a:06x}  {t}")
