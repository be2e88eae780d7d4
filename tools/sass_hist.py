#!/usr/bin/env python
"""SASS instruction histogram of kernels in a cubin / shared object (cuobjdump -sass), for profiles/.

    python tools/sass_hist.py parallel_finite_difference_computation_b200/libfdwave.so 'k_step<8, 0, 0>|k_lap<8>'
"""
import collections
import re
import subprocess
import sys


def main():
    so, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if not re.search(pat, d):
            continue
        ops = collections.Counter()
        for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", f):
            op = m.group(1)
            if op.split(".")[0] == "NOP":
                continue
            key = op.split(".")[0]
            if key in ("LDG", "STG", "LDS", "STS", "LD", "ST", "LDL", "STL", "F2F", "ATOMG", "RED", "LDGSTS", "UBLKCP", "UTMALDG"):
                key = ".".join(op.split(".")[:3])
            ops[key] += 1
        print("%s\n  static instructions: %d" % (d, sum(ops.values())))
        print("  " + "  ".join("%s:%d" % kv for kv in ops.most_common()))


if __name__ == "__main__":
    main()
