#!/usr/bin/env python3
"""Re-derive nquant_android_b200/csrc/nq_bluenoise_table.h from the reference's data table
(BlueNoise.java:13-178). Only runs where /root/reference exists; the generated header is committed.
sha256 of the 4096 bytes: see tests/test_oracle.py::test_blue_noise_table."""
import re, sys
src = open('/root/reference/nQuant.master/src/main/java/com/android/nQuant/BlueNoise.java').read()
body = src[src.index('TELL_BLUE_NOISE = {') + len('TELL_BLUE_NOISE = {'):]
body = body[:body.index('};')]
vals = [int(v) for v in re.findall(r'-?\d+', body)]
assert len(vals) == 4096
lines = ["// 64x64 scalar blue-noise mask (Tellusim 64x64_l64_s16, signed bytes), the data table the reference",
         "// indexes as TELL_BLUE_NOISE (BlueNoise.java:13-178). Data only; emitted by tools/extract_bluenoise.py.",
         "#pragma once", "#define NQ_BLUE_NOISE_INIT { \\"]
for i in range(0, 4096, 32):
    lines.append("  " + ",".join(str(v) for v in vals[i:i + 32]) + ", \\")
lines.append("}")
open('nquant_android_b200/csrc/nq_bluenoise_table.h', 'w').write("\n".join(lines) + "\n")
