"""Prints the 256-entry sRGB -> linear table of vknrc_b200/csrc/nrc_unpack.cuh (formula in float64, rounded to float32)."""
import numpy as np

vals = []
for c in range(256):
    x = c / 255.0
    vals.append(np.float32(x / 12.92 if x <= 0.04045 else ((x + 0.055) / 1.055) ** 2.4))
for i in range(0, 256, 8):
    print("\t" + ", ".join(("%.9g" % float(v) + ("" if "." in "%.9g" % float(v) or "e" in "%.9g" % float(v) else ".0") + "f") for v in vals[i:i + 8]) + ",")
