#!/usr/bin/env python
"""Regenerates the 32-entry table of include/mkf_expf.h: T[i] = bits(2^(i/32)) - (i << 47), with 2^(i/32)
correctly rounded to double (mpmath at 200 bits), and checks it against the header."""
import os
import re
import struct

import mpmath

mpmath.mp.prec = 200
tab = []
for i in range(32):
    d = float(mpmath.power(2, mpmath.mpf(i) / 32))
    tab.append(struct.unpack("<Q", struct.pack("<d", d))[0] - (i << 47))
hdr = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "mkf_expf.h")).read()
body = hdr[hdr.index("#define MKF_EXPF_TABLE"):hdr.index("static const uint64_t mkf_expf_T_host")]
have = [int(x, 16) for x in re.findall(r"0x[0-9a-f]{16}", body)]
print(", ".join(hex(t) for t in tab))
assert have == tab, "include/mkf_expf.h table differs from the regenerated one"
print("table in include/mkf_expf.h matches")
